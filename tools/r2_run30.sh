#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for u in 16 8 4 32; do echo "== LDB200_UNITS_PER_SM=$u"; LDB200_UNITS_PER_SM=$u timeout 600 python tools/ab_rigid.py 2>&1 | tail -3; done > gpurun_out/r2_run30_units_flex.log 2>&1
cat gpurun_out/r2_run30_units_flex.log
