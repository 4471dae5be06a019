#!/bin/bash
cd "$(dirname "$0")/.."
TAG=${1:-r2_dna}
ncu --set full --clock-control none --import-source on -k regex:dna_pair_kernel -s 2 -c 1 -f -o gpurun_out/prof_$TAG \
    python tools/dna_timing.py > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log | cut -c1-200
