#!/bin/bash
# ncu --set full captures of the generic DFIRE kernel (2uuy, ligand ANM) and the DNA kernel (1azp) on the
# 20,000-pose launch of tools/config_timings.py (the 53rd launch of a case is its timed large batch).
ncu --set full --clock-control none --import-source on -k regex:dfire_pair_kernel -s 105 -c 1 -f -o gpurun_out/prof_r1_dfire_pair_kernel \
    python tools/config_timings.py > gpurun_out/ncu_r1_dfire_pair_kernel.log 2>&1
tail -2 gpurun_out/ncu_r1_dfire_pair_kernel.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:dna_pair_kernel -s 52 -c 1 -f -o gpurun_out/prof_r1_dna_pair_kernel \
    python tools/config_timings.py > gpurun_out/ncu_r1_dna_pair_kernel.log 2>&1
tail -2 gpurun_out/ncu_r1_dna_pair_kernel.log | cut -c1-200
