#!/bin/bash
# ncu counters of the dominant kernel of every configuration bench.py reports (run under gpurun, 1 GPU).  One launch
# each: the first launch of that kernel inside the last (third) call of tools/count_target.py (--profile-from-start off).
# Output: gpurun_out/counts_<config>.csv; tools/make_counts.py turns them into profiles/r2_kernel_counts.json.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_op_read.sum,launch__registers_per_thread,launch__block_size,launch__grid_size
for spec in 1k4c_bench:dfire_rigid_kernel 1ppe:dfire_rigid_kernel 1czy:dfire_rigid_kernel 2uuy:dfire_rigid_kernel ab_icode:dfire_rigid_kernel 1azp:dna_pair_kernel; do
  cfg=${spec%%:*}; k=${spec#*:}
  ncu --metrics $M --clock-control none -k regex:$k --profile-from-start off -c 1 --csv --log-file gpurun_out/counts_$cfg.csv \
      python tools/count_target.py $cfg > gpurun_out/counts_$cfg.log 2>&1
  tail -1 gpurun_out/counts_$cfg.log | cut -c1-160
done
