#!/bin/bash
# round-2 GPU run 24: final single-GPU measurement set: kernel counters, launch list, default bench line, DNA kernel capture
cd "$(dirname "$0")/.."
bash tools/r2_run9.sh
ncu --set full --clock-control none --import-source on -k regex:dna_pair_kernel --profile-from-start off -c 1 -o gpurun_out/prof_r2f_1azp_dna -f \
    python tools/count_target.py 1azp > gpurun_out/ncu_r2f_1azp_dna.log 2>&1
tail -1 gpurun_out/ncu_r2f_1azp_dna.log | cut -c1-160
