#!/bin/bash
# ncu --set full (+ source) captures of the shipped kernels, steady-state launches: rigid (1k4c bench step, 1ppe), FLEX (2uuy),
# DNA (1azp)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cap() {  # tag config kernel-regex
  ncu --set full --clock-control none --import-source on -k regex:$3 --profile-from-start off -c 1 -o gpurun_out/prof_r2f_$1 -f \
      python tools/count_target.py $2 > gpurun_out/ncu_r2f_$1.log 2>&1
  tail -1 gpurun_out/ncu_r2f_$1.log | cut -c1-160
}
cap 1k4c_bench 1k4c_bench dfire_rigid_kernel
cap 1ppe 1ppe dfire_rigid_kernel
cap 2uuy_flex 2uuy dfire_rigid_kernel
cap 1azp_dna 1azp dna_pair_kernel
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
