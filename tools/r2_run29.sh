#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:dfire_rigid_kernel --profile-from-start off -c 1 -o gpurun_out/prof_r2g_2uuy_flex -f \
    python tools/count_target.py 2uuy > gpurun_out/ncu_r2g_2uuy_flex.log 2>&1
tail -1 gpurun_out/ncu_r2g_2uuy_flex.log | cut -c1-160
