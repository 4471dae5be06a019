#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for spec in 1024:16 2048:16 4096:16 8192:16 4096:8 8192:8 8192:4; do m=${spec%%:*}; u=${spec#*:}; echo "== LDB200_UNIT_MAX_POSES=$m LDB200_UNITS_PER_SM=$u"; LDB200_UNIT_MAX_POSES=$m LDB200_UNITS_PER_SM=$u timeout 600 python tools/ab_rigid.py 2>&1 | grep -E "1k4c (10000|80000)|rank 0|1ppe"; done > gpurun_out/r2_run47_units.log 2>&1
cat gpurun_out/r2_run47_units.log
