#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rigid_path.py -m gpu -x -q -k "interface_pass" 2>&1 | tail -3
for c in 0.75 0.6; do echo "== LDB200_CELL=$c"; LDB200_CELL=$c timeout 600 python tools/ab_rigid.py 2>&1 | grep -E "poses|parity"; done > gpurun_out/r2_run37_cell.log 2>&1
cat gpurun_out/r2_run37_cell.log
