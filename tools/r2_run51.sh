#!/bin/bash
# round-2 GPU run 51: --set full capture of the FLEX instance on 1czy (final build)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:dfire_rigid_kernel --profile-from-start off -c 1 -o gpurun_out/prof_r2f_1czy_flex -f \
    python tools/count_target.py 1czy > gpurun_out/ncu_r2f_1czy_flex.log 2>&1
tail -1 gpurun_out/ncu_r2f_1czy_flex.log | cut -c1-160
