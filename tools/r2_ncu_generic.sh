#!/bin/bash
# ncu capture of the generic DFIRE kernel on the 20,000-pose 2uuy batch of tools/config_profile.py (3rd case; each case
# launches the pair kernel 3 times per path: warm-up, timed, detail)
cd "$(dirname "$0")/.."
TAG=${1:-r2_generic}
ncu --set full --clock-control none --import-source on -k regex:dfire_pair_kernel -s 4 -c 1 -f -o gpurun_out/prof_$TAG \
    python tools/config_profile.py > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log | cut -c1-200
