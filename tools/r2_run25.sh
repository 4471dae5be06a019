#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_device_gso.py -m gpu -x -q 2>&1 | tail -5
python tools/gso_timing.py 2>&1 | tail -6
