#!/bin/bash
# round-2 GPU run 7: pose-major kernel: tests, then A/B (LDB200_POSE_MAJOR=0 vs auto) timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_rigid_path.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2_run7_pytest.log
tail -30 gpurun_out/r2_run7_pytest.log
for pm in 0 -1; do
  export LDB200_POSE_MAJOR=$pm
  echo "== LDB200_POSE_MAJOR=$pm"
  timeout 600 python tools/ab_rigid.py 2>&1 | tail -4
  timeout 600 python tools/config_timings.py 2>&1 | tail -12
done > gpurun_out/r2_run7_ab.log 2>&1
cat gpurun_out/r2_run7_ab.log
