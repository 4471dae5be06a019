#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in 0.7 0.8 0.85 0.9; do echo "== LDB200_CELL=$c"; LDB200_CELL=$c timeout 600 python tools/ab_rigid.py 2>&1 | grep -E "1k4c (10000|80000)|rank 0"; done > gpurun_out/r2_run41_cell.log 2>&1
cat gpurun_out/r2_run41_cell.log
