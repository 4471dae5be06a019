#!/bin/bash
# round-2 GPU run 22: DNA/pyDock pose transform fused into the pair kernel: tests + A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trajectory.py tests/test_gpu_device_gso.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2_run22_pytest.log
tail -12 gpurun_out/r2_run22_pytest.log
for f in 0 1; do echo "== LDB200_DNA_FUSED=$f"; LDB200_DNA_FUSED=$f timeout 300 python tools/dna_timing.py 2>&1 | grep -E "parity|poses " | tail -3; LDB200_DNA_FUSED=$f timeout 600 python tools/config_timings.py 2>&1 | grep 1azp; done > gpurun_out/r2_run22_dna.log 2>&1
cat gpurun_out/r2_run22_dna.log
