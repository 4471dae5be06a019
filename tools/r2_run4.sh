#!/bin/bash
# round-2 GPU run 4: device-resident GSO tests, then the whole GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_device_gso.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_run4_gso.log
tail -40 gpurun_out/r2_run4_gso.log
timeout 1800 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_device_gso.py 2>&1 | tail -8 > gpurun_out/r2_run4_pytest.log
tail -8 gpurun_out/r2_run4_pytest.log
