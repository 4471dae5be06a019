#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for u in 16 8 12 24 32; do echo "== LDB200_UNITS_PER_SM=$u"; LDB200_UNITS_PER_SM=$u timeout 600 python tools/ab_rigid.py 2>&1 | grep -E "1ppe|2uuy|1czy"; done > gpurun_out/r2_run58_units_flex.log 2>&1
cat gpurun_out/r2_run58_units_flex.log
