#!/bin/bash
# round-2 GPU run 8: ncu --set full (+ source) of the ligand-frame kernel on 2uuy (FLEX), 1ppe and the bench step,
# group-major and pose-major instances
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cap() {  # tag config kernel-regex env
  env $4 ncu --set full --clock-control none --import-source on -k regex:$3 -s 1 -c 1 -o gpurun_out/prof_r2_$1 -f \
      python tools/count_target.py $2 > gpurun_out/ncu_r2_$1.log 2>&1
  tail -2 gpurun_out/ncu_r2_$1.log | cut -c1-200
}
cap 2uuy_gm 2uuy dfire_rigid_kernel LDB200_POSE_MAJOR=0
cap 2uuy_pm 2uuy dfire_posemajor_kernel LDB200_POSE_MAJOR=1
cap 1ppe_gm 1ppe dfire_rigid_kernel LDB200_POSE_MAJOR=0
cap 1ppe_pm 1ppe dfire_posemajor_kernel LDB200_POSE_MAJOR=1
cap 1k4c_bench 1k4c_bench dfire_rigid_kernel LDB200_POSE_MAJOR=0
cap 1azp_dna 1azp dna_pair_kernel LDB200_POSE_MAJOR=0
ls -la gpurun_out/prof_r2_*.ncu-rep
