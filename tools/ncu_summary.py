#!/usr/bin/env python3
"""Summarise an .ncu-rep (read with `ncu -i`) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "sm__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        print("== kernel:", d.get("Kernel Name", "?")[:80])
        for k in KEYS:
            if k in d: print(f"{k:80s} {d[k]:>18s} {u[k]}")
        print("-- warp stall reasons (avg warps stalled per issue-active cycle)")
        st = [(k, float(v)) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and v]
        for k, v in sorted(st, key=lambda x: -x[1])[:9]:
            print(f"   {k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):28s} {v:8.3f}")
if __name__ == "__main__":
    main(sys.argv[1])
