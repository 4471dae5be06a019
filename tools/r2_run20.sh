#!/bin/bash
# round-2 GPU run 20: ligand-frame cell size with the compact tiles
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in 1.0 0.8 0.7 0.6 1.25; do echo "== LDB200_CELL=$c"; LDB200_CELL=$c timeout 600 python tools/ab_rigid.py 2>&1 | tail -5; done > gpurun_out/r2_run20_cell.log 2>&1
cat gpurun_out/r2_run20_cell.log
