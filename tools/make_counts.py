#!/usr/bin/env python3
"""gpurun_out/counts_<config>.csv (tools/capture_counts.sh) -> profiles/r2_kernel_counts.json, stamped with the hash
of the kernel sources so that bench.py can refuse counters that belong to another build."""
import csv
import glob
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def source_sha():
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, "lightdock-rust_b200", "csrc", "*"))):
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()


def main():
    out = {"source_sha256": source_sha(), "how": "ncu --metrics ... --clock-control none, tools/capture_counts.sh: the "
           "third ld_score_batch call of tools/count_target.py <config> (20,000 poses; 1k4c_bench: 60,000 poses of "
           "the bench step)", "kernels": {}}
    poses = {"1k4c_bench": 60000}
    for path in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "counts_*.csv"))):
        cfg = os.path.basename(path)[len("counts_"):-4]
        rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
        hdr = rows[0]
        vals = {}
        for r in rows[1:]:
            d = dict(zip(hdr, r))
            vals[d["Metric Name"]] = (float(d["Metric Value"].replace(",", "")), d["Metric Unit"])
            kname = d["Kernel Name"]
        if not vals:
            continue
        # poses of the captured launch: printed by tools/count_target.py ("<config> <n> poses ...")
        n_poses = poses.get(cfg, 20000)
        log = os.path.join(ROOT, "gpurun_out", f"counts_{cfg}.log")
        if os.path.exists(log):
            for line in open(log, errors="ignore"):
                w = line.split()
                if len(w) >= 3 and w[0] == cfg and w[2] == "poses" and w[1].isdigit():
                    n_poses = int(w[1])
        def v(name, scale=None):
            x, unit = vals[name]
            if scale == "bytes":
                x *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
            if scale == "ms":
                x *= {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}[unit]
            return x
        out["kernels"][cfg] = {
            "kernel": kname.split("(")[0], "poses": n_poses,
            "warp_inst": v("smsp__inst_executed.sum"), "thread_inst": v("smsp__thread_inst_executed.sum"),
            "dram_read_bytes": v("dram__bytes_read.sum", "bytes"), "dram_write_bytes": v("dram__bytes_write.sum", "bytes"),
            "smem_wavefronts": v("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
            "smem_bank_conflict_wavefronts": v("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
            "l2_read_sectors": v("lts__t_sectors_srcunit_tex_op_read.sum"),
            "issue_active_pct_under_ncu": v("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "fp64_pipe_active_pct_under_ncu": v("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "warps_active_pct_under_ncu": v("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "duration_ms_under_ncu": v("gpu__time_duration.sum", "ms"),
            "registers": v("launch__registers_per_thread"), "block_size": v("launch__block_size"), "grid_size": v("launch__grid_size")}
    path = os.path.join(ROOT, "profiles", "r2_kernel_counts.json")
    json.dump(out, open(path, "w"), indent=1)
    print(path, list(out["kernels"]))


if __name__ == "__main__":
    main()
