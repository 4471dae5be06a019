#!/usr/bin/env python3
"""bench.py — headline benchmark of the LightDock scoring hot path on B200.

Metric (BASELINE.json): glowworm poses scored / s (and atom-pair evals / s) on the synthetic 1k4c-shaped workload
(configs[4]): the real 1k4c membrane receptor (3413 atoms incl. 453 MMB beads) and ligand (3268 atoms), DFIRE scoring,
400 swarms x 200 glowworms, swarms sharded over the GPUs by the cost-aware deterministic map of host/sharding.hpp.
A "step" is one GSO-step scoring pass: every pose of every swarm owned by the rank is scored by one ld_score_batch call.
DFIRE uses the seeded synthetic DCparams unless $LIGHTDOCK_DATA holds a real one.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

`value`   : poses/s with the pose rows already resident in HBM (CUDA events on torch's current stream, the stream the
            kernels are launched on), whole job = sum over ranks / max time over ranks.
`e2e`     : the same through the reference-facing C ABI with HOST buffers (ld_score_batch: pinned staging, H2D of the pose
            rows, kernels, D2H of the energies inside the timed region).
`roofline`: the dominant kernel (dfire_rigid_kernel) against what actually bounds it: warp-instruction issue (4 schedulers x
            SMs x SM clock), with the shared-memory wavefront rate as co-limiter; instruction / wavefront / DRAM counts per
            pose come from profiles/r2_kernel_counts.json (ncu capture of the same kernel), which carries the hash of the
            kernel sources: counters of another build are refused.  SURVEY.md §8(d)'s figure (8 FP64 flops per brute-force
            pair against the measured non-fused FP64 rate) is kept as `bruteforce_equivalent`: the kernel prunes 98 % of
            the pairs and classifies in FP32, so that ratio measures pruning, not how close to a limit the kernel runs.
`configs` : (N=1) BASELINE configs 0-3 -- 1czy, 1ppe, 2uuy (the shipped DFIRE + ANM protein-protein example), 1azp (DNA):
            e2e poses/s on a 20,000-pose batch through ld_score_batch, the dominant kernel's time and roofline fraction,
            and the CPU port on the same poses.
`cpu_baseline`: the oracle port (oracle/ld_oracle.c, same scalar f64 loop, -O2 without -march=native like a default
            `cargo build --release`) on all host cores -- threads of ONE process over independent poses, equivalent to the
            reference's one-process-per-core model because poses share nothing -- bounded sample.
`--impl reference`: only that CPU arm (the Rust reference cannot be built here: no cargo).
"""
import argparse
import glob
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "lightdock-rust_b200"))

N_REC, N_LIG = 3413, 3268
METRIC, UNIT = "glowworm_poses_scored_per_s", "poses/s"
COUNTS_FILE = os.path.join(ROOT, "profiles", "r2_kernel_counts.json")


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def workload_config(args, world):
    return {"workload": "synthetic 1k4c-sized: real 1k4c receptor (3413 atoms, 453 MMB beads) + ligand (3268 atoms), "
                        f"DFIRE, {args.swarms} swarms x {args.glowworms} glowworms, every pose rescored each step",
            "swarms": args.swarms, "glowworms": args.glowworms, "poses_per_step": args.swarms * args.glowworms,
            "pairs_per_pose": N_REC * N_LIG,
            "sharding": f"cost-aware deterministic map over {world} GPU(s) (expected in-reach pairs at each swarm's "
                        "centre, longest-processing-time first; host/sharding.hpp), no collective on the data path",
            "l2": "flushed between steps (256 MiB memset inside the timed region)"}


def kernel_source_sha():
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, "lightdock-rust_b200", "csrc", "*"))):
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()


def load_counts():
    """Per-kernel ncu counters of THIS build, or (None, why)."""
    try:
        c = json.load(open(COUNTS_FILE))
    except OSError:
        return None, "profiles/r2_kernel_counts.json missing"
    if c.get("source_sha256") != kernel_source_sha():
        return None, "profiles/r2_kernel_counts.json was captured from other kernel sources (hash mismatch): refused"
    return c["kernels"], None


# ------------------------------------------------------------------------------------------------
def oracle_setup():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as O
    return O


def oracle_complex():
    """CPU baseline leg only: the oracle's model of the same complex."""
    O = oracle_setup()
    from ldb200 import workload
    pot = workload.synthetic_dcparams()
    d = os.environ.get("LIGHTDOCK_DATA")
    if d and os.path.exists(os.path.join(d, "DCparams")):
        pot = O.load_dcparams(os.path.join(d, "DCparams"))
    cx, _, _, _ = O.load_case(os.path.join(ROOT, "tests", "golden", "1k4c"), O.DFIRE, potential=pot)
    return cx


def cpu_rate(cx, poses, cores, target_s):
    """poses/s of the oracle on `cores` threads over a bounded sample of ~target_s seconds."""
    n0 = min(len(poses), cores * 4)
    t = time.perf_counter()
    cx.energy_mt(poses[:n0], cores)
    dt = time.perf_counter() - t
    n = int(min(len(poses), max(cores * 4, (n0 / dt) * target_s)))
    n -= n % cores
    n = max(n, cores)
    t = time.perf_counter()
    cx.energy_mt(poses[:n], cores)
    dt = time.perf_counter() - t
    return n / dt, n, dt


CPU_NOTE = ("oracle/ld_oracle.c port of the reference's scalar f64 loop (gcc -O2 -ffp-contract=off, no -march=native: what "
            "a default `cargo build --release` emits), threads of one process over independent poses = the reference's "
            "one-process-per-core fan-out (poses share nothing); the Rust reference itself cannot be built in this image "
            "(no cargo/rustc)")


def run_reference(args, rank, world):
    """--impl reference: the reference algorithm's CPU implementation (oracle port) on all host cores."""
    if rank != 0:
        return
    from ldb200 import workload
    cores = host_cores()
    cx = oracle_complex()
    poses = workload.synthetic_1k4c_swarms(args.swarms, args.glowworms).reshape(-1, 7)
    rng = np.random.default_rng(1)
    poses = poses[rng.permutation(len(poses))]
    per_step = cores * 24  # bounded sample of the workload per step (~0.3 s of work per core)
    for w in range(args.warmup):
        cx.energy_mt(poses[w * per_step:(w + 1) * per_step], cores)
    t0 = time.perf_counter()
    for k in range(args.steps):
        o = ((args.warmup + k) * per_step) % (len(poses) - per_step)
        cx.energy_mt(poses[o:o + per_step], cores)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = f"{per_step} poses/step ({cores} threads x 24) drawn from the same {len(poses)}-pose workload"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "pair_evals_per_s": value * N_REC * N_LIG,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": CPU_NOTE}))


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def summary(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7 or not (t0 <= ts <= t1):
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def issue_roofline(cnt, poses, kernel_ms, sm_hz, sms):
    """Warp-instruction issue: achieved / (4 schedulers x SMs x clock), and the shared-memory wavefront rate beside it."""
    inst = cnt["warp_inst"] / cnt["poses"] * poses
    wf = cnt["smem_wavefronts"] / cnt["poses"] * poses
    sec = kernel_ms * 1e-3
    return {"achieved_gwarp_inst_per_s": inst / sec / 1e9, "peak_gwarp_inst_per_s": sms * 4 * sm_hz / 1e9,
            "frac": inst / sec / (sms * 4 * sm_hz), "warp_inst_per_pose": cnt["warp_inst"] / cnt["poses"],
            "smem_wavefront_frac": wf / sec / (sms * sm_hz),
            "smem_wavefronts_per_pose": cnt["smem_wavefronts"] / cnt["poses"],
            "smem_bank_conflict_share": cnt["smem_bank_conflict_wavefronts"] / max(1.0, cnt["smem_wavefronts"]),
            "avg_active_threads": cnt["thread_inst"] / max(1.0, cnt["warp_inst"]),
            "issue_active_pct_under_ncu": cnt["issue_active_pct_under_ncu"]}


# ------------------------------------------------------------------------------------------------
def config_batch(name, pose_len):
    """The 20,000-pose batch of a BASELINE config: its 200 shipped start poses x 100, translations jittered by 1 A
    (seeded), ANM extents kept (the same batch tools/count_target.py captures the kernel counters on)."""
    g = os.path.join(ROOT, "tests", "golden", name)
    start = os.path.join(g, "initial_positions_0.dat")
    if not os.path.exists(start):
        start = os.path.join(g, "init", "initial_positions_0.dat")
    pos = np.loadtxt(start)[:, :pose_len]
    rng = np.random.default_rng(1)
    big = np.tile(pos, (100, 1))
    big[:, :3] += rng.normal(0, 1.0, size=(len(big), 3))
    return g, np.ascontiguousarray(big)


def config_leg(device, peaks_fp64_tf, sm_hz, sms, counts, cores, cpu_seconds):
    """BASELINE configs 0-3 on one GPU: 20,000-pose batches through the product host layer (setup.json -> CudaScore ->
    ld_score_batch with host buffers).  The oracle appears only as the CPU baseline / checker of each config."""
    import ldb200
    from ldb200 import host
    lib = ldb200.load_library()
    out = {}
    for name, method in (("1czy", "dfire"), ("1ppe", "dfire"), ("2uuy", "dfire"), ("1azp", "dna")):
        g = os.path.join(ROOT, "tests", "golden", name)
        c = host.Case(os.path.join(g, "setup.json"), method, anm_dir=g, device=device)
        h = c.ld_handle()
        _, poses = config_batch(name, c.pose_len)
        n_pairs = c.n_rec * c.n_lig
        c.energy_batch(poses)          # warm-up (FLEX handles learn their slacks and rebuild their lists here)
        c.energy_batch(poses)
        reps = 5
        t = time.perf_counter()
        for _ in range(reps):
            e = c.energy_batch(poses)
        dt = (time.perf_counter() - t) / reps
        lib.ld_set_profiling(h, 1)
        c.energy_batch(poses)
        st = ldb200.handle_stats(lib, h)
        lib.ld_set_profiling(h, 0)
        smp = np.ascontiguousarray(poses[:500])
        det = (ldb200.PoseDetail * len(smp))()
        es = np.empty(len(smp))
        ldb200._check(lib, lib.ld_score_batch_detail(h, len(smp), smp.ctypes.data, es.ctypes.data, det, None, None))
        dna = method == "dna"
        kernel = "dna_pair_kernel" if dna else ("dfire_rigid_kernel" if st["path"] == ldb200.PATH_RIGID else "dfire_pair_kernel")
        roof = {"kernel": kernel, "kernel_ms": st["pair_ms"], "prep_or_transform_ms": st["transform_ms"],
                "finalize_ms": st["finalize_ms"]}
        if dna:
            # 10.4 algorithmic flops per brute-force pair against the measured non-fused FP64 rate
            ach = 10.4 * n_pairs * len(poses) / (st["pair_ms"] * 1e-3) / 1e12
            roof.update({"bound": "fp64", "achieved": ach, "peak": peaks_fp64_tf, "unit": "TFLOP/s", "frac": ach / peaks_fp64_tf,
                         "algorithmic_flops_per_pair": 10.4,
                         "note": "the kernel uses FMAs (9 FP64 instructions per far pair, 15 per close pair), so the "
                                 "fraction of the NON-fused probe can pass 1"})
        else:
            ach = 8.0 * n_pairs * len(poses) / (st["pair_ms"] * 1e-3) / 1e12
            roof.update({"bound": "issue", "achieved": None, "peak": sms * 4 * sm_hz / 1e9, "unit": "Gwarp-inst/s", "frac": None,
                         "bruteforce_equivalent": {"tflops": ach, "vs_fp64_probe": ach / peaks_fp64_tf,
                                                   "note": "pruning factor, not a roofline fraction"}})
        if counts and name in counts and counts[name]["kernel"].split("<")[0].endswith(kernel):
            roof["issue"] = issue_roofline(counts[name], len(poses), st["pair_ms"], sm_hz, sms)
            if not dna:
                roof.update({"achieved": roof["issue"]["achieved_gwarp_inst_per_s"], "frac": roof["issue"]["frac"]})
            roof["traffic"] = (counts[name]["dram_read_bytes"] + counts[name]["dram_write_bytes"]) / counts[name]["poses"] * len(poses)
        # CPU baseline + check: the oracle's model of the same complex on the same poses
        O = oracle_setup()
        pot = None
        if not dna:
            pot = O.load_dcparams(os.path.join(os.environ["LIGHTDOCK_DATA"], "DCparams"))
        cx, _, _, _ = O.load_case(g, O.DNA if dna else O.DFIRE, potential=pot)
        rate, n_cpu, cdt = cpu_rate(cx, poses, cores, cpu_seconds)
        e_cpu = cx.energy(poses[:8])
        out[name] = {
            "method": method.upper(), "n_rec": c.n_rec, "n_lig": c.n_lig, "pairs_per_pose": n_pairs,
            "pose_len": c.pose_len, "poses": len(poses),
            "e2e": {"value": len(poses) / dt, "unit": UNIT, "ms_per_batch": dt * 1e3, "h2d_bytes": poses.size * 8,
                    "d2h_bytes": len(poses) * 8, "api": "CudaScore::energy_batch -> ld_score_batch (C ABI, host buffers)"},
            "pair_evals_per_s": len(poses) / dt * n_pairs, "path_info": c.path_info()[:200], "roofline": roof,
            "executed_pair_test_fraction": float(np.mean([d.n_pairs_tested for d in det]) / n_pairs),
            "in_cutoff_fraction": float(np.mean([d.n_in_cutoff for d in det]) / n_pairs),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n_cpu} of the same poses in {cdt:.1f} s on {cores} threads"},
            "gpu_vs_oracle_max_rel_err_8_poses": float(np.max(np.abs(e_cpu - e[:8]) / np.abs(e_cpu)))}
        c.close()
    return out


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import ldb200
    from ldb200 import host, workload

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("WARN", "VERSION"):
            os.environ.pop("NCCL_DEBUG")  # those levels printf "NCCL version ..." on stdout, next to the JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = ldb200.load_library()

    dc_dir, table_kind = workload.ensure_dcparams_dir(tempfile.gettempdir())
    os.environ["LIGHTDOCK_DATA"] = dc_dir
    case = host.Case(os.path.join(workload.GOLDEN_1K4C, "setup.json"), "dfire", device=local_rank)
    assert (case.n_rec, case.n_lig, case.pose_len) == (N_REC, N_LIG, 7)
    h = case.ld_handle()

    all_poses = workload.synthetic_1k4c_swarms(args.swarms, args.glowworms)
    rec_xyz = workload.read_pdb_coords(os.path.join(workload.GOLDEN_1K4C, "lightdock_receptor_membrane.pdb"))
    lig_xyz = workload.read_pdb_coords(os.path.join(workload.GOLDEN_1K4C, "lightdock_ligand.pdb"))
    mine = workload.shard_swarms_cost_aware(all_poses, rec_xyz, lig_xyz, rank, world)
    poses = np.ascontiguousarray(all_poses[mine].reshape(-1, 7))
    n_local = poses.shape[0]
    n_total = args.swarms * args.glowworms
    dev = torch.device("cuda", local_rank)
    d_poses = torch.from_numpy(poses).to(dev)
    d_energy = torch.zeros(n_local, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # a real (non-NULL) stream: the C ABI reads a NULL stream as "use the handle's own stream", and the
    # CUDA events below must sit on the stream the kernels are launched on
    work_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(work_stream)
    stream = work_stream.cuda_stream
    assert stream != 0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def per_rank(x):
        if world == 1:
            return [x]
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = x
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.cpu()]

    def step_device():
        flush.zero_()
        ldb200._check(lib, lib.ld_score_batch_device(h, n_local, d_poses.data_ptr(), d_energy.data_ptr(), stream))

    # ---- device-resident throughput (`value`) --------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.3)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches_per_step = ldb200.handle_stats(lib, h)["kernel_launches"]
    value = n_total * args.steps / (ms_total * 1e-3)
    e_dev = d_energy.cpu().numpy()

    # ---- per-kernel time of the same step (events inside the library, same stream) -------------
    lib.ld_set_profiling(h, 1)
    pair_ms = tr_ms = fin_ms = 0.0
    for _ in range(args.steps):
        step_device()
        st = ldb200.handle_stats(lib, h)
        pair_ms += st["pair_ms"]; tr_ms += st["transform_ms"]; fin_ms += st["finalize_ms"]
    lib.ld_set_profiling(h, 0)
    st_last = ldb200.handle_stats(lib, h)
    pair_launches = st_last["pair_launches"] * args.steps
    path = {ldb200.PATH_GENERIC: "generic", ldb200.PATH_RIGID: "rigid"}[st_last["path"]]
    pair_kernel = "dfire_rigid_kernel" if path == "rigid" else "dfire_pair_kernel"
    pair_ms_step = max_over_ranks(pair_ms / args.steps)
    pair_ms_ranks = per_rank(pair_ms / args.steps)
    poses_ranks = per_rank(float(n_local))

    # ---- end to end through the C ABI with host buffers (`e2e`) --------------------------------
    e_host = np.empty(n_local)
    for _ in range(min(args.warmup, 3)):
        ldb200._check(lib, lib.ld_score_batch(h, n_local, poses.ctypes.data, e_host.ctypes.data))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        ldb200._check(lib, lib.ld_score_batch(h, n_local, poses.ctypes.data, e_host.ctypes.data))
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    t_wall2 = time.perf_counter()
    barrier()
    # clocks sampled over every timed leg of the step (device-resident, per-kernel, host-buffer): all under load
    clk = clocks.summary(t_wall0, t_wall2) if clocks else None
    if clk is not None:
        clk["window"] = "value + per-kernel + e2e legs (%.0f ms), nvidia-smi every 20 ms" % ((t_wall2 - t_wall0) * 1e3)
    e2e_value = n_total * args.steps / e2e_s
    if not np.array_equal(e_host, e_dev):
        raise SystemExit("bench.py: host-buffer and device-buffer paths disagree")

    # ---- whole GSO run through the host driver (MultiGSO: real control flow, RNG stream, moved-only rescoring) ----
    gso = None
    if args.gso_steps > 0:
        threads = max(1, min(32, host_cores() // world))
        seeds = np.full(len(mine), 324324, dtype=np.uint64)  # every swarm is seeded like a stand-alone reference process
        case.multi_gso(all_poses[mine], seeds, 1, host_threads=threads)  # warm-up: sizes both slots' buffers
        barrier()
        t0 = time.perf_counter()
        final_state, calls = case.multi_gso(all_poses[mine], seeds, args.gso_steps, host_threads=threads)
        gso_s = max_over_ranks(time.perf_counter() - t0)
        calls_t = torch.tensor([calls], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(calls_t, op=dist.ReduceOp.SUM)
        gso = {"steps": args.gso_steps, "swarms": args.swarms, "energy_calls": int(calls_t.item()),
               "wall_s": gso_s, "poses_per_s": calls_t.item() / gso_s, "host_threads_per_rank": threads,
               "moved_fraction": calls_t.item() / (n_total * args.gso_steps),
               "what": "lightdock host GSO (C++ MultiGSO): real control flow and RNG streams, moved-only rescoring, two swarm sets "
                       "leapfrogging over ld_score_batch_begin/_end; wall clock incl. every host phase"}
        if rank == 0:
            # why this leg can run above `value`: the swarms drift (away from the receptor on the synthetic table), so
            # a pose of step k has fewer in-reach pairs than a start pose.  Executed pair tests per pose, sampled.
            sample0 = poses[:: max(1, n_local // 256)][:256]
            sample1 = np.ascontiguousarray(final_state.reshape(-1, final_state.shape[-1])[:: max(1, n_local // 256), 4:][:256])
            tested = []
            for smp in (sample0, sample1):
                det = (ldb200.PoseDetail * len(smp))()
                es = np.empty(len(smp))
                ldb200._check(lib, lib.ld_score_batch_detail(h, len(smp), np.ascontiguousarray(smp).ctypes.data,
                                                             es.ctypes.data, det, None, None))
                tested.append(float(np.mean([d.n_pairs_tested for d in det])))
            gso["pair_tests_per_pose_start"] = tested[0]
            gso["pair_tests_per_pose_after_gso"] = tested[1]
            gso["note"] = ("the kernel's time follows the executed pair tests: poses after %d GSO steps need %.2fx the pair "
                           "tests of the start poses `value` is measured on" % (args.gso_steps, tested[1] / max(tested[0], 1.0)))

    # ---- the same GSO run with the whole step on the device (DeviceGSO -> ld_gso_*) ----
    gso_dev = None
    if args.gso_steps > 0:
        case.device_gso(all_poses[mine], seeds, 1, host_threads=threads)  # warm-up
        barrier()
        t0 = time.perf_counter()
        dev_state, dev_calls = case.device_gso(all_poses[mine], seeds, args.gso_steps, host_threads=threads)
        dev_s = max_over_ranks(time.perf_counter() - t0)
        calls_d = torch.tensor([dev_calls], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(calls_d, op=dist.ReduceOp.SUM)
        same_discrete = bool(np.array_equal(dev_state[..., 2:4], final_state[..., 2:4]) and dev_calls == calls)
        gso_dev = {"steps": args.gso_steps, "swarms": args.swarms, "energy_calls": int(calls_d.item()), "wall_s": dev_s,
                   "poses_per_s": calls_d.item() / dev_s,
                   "vs_host_loop": {"neighbour_counts_vision_ranges_energy_calls_identical": same_discrete,
                                    "max_abs_pose_diff": float(np.abs(dev_state[..., 4:] - final_state[..., 4:]).max()),
                                    "max_rel_scoring_diff": float(np.max(np.abs(dev_state[..., 1] - final_state[..., 1])
                                                                         / np.maximum(np.abs(final_state[..., 1]), 1e-300)))},
                   "what": "the same swarms, seeds and steps with the GSO step itself on the GPU (gso_step_kernel + moved-only "
                           "scoring pass per step, no host round trip; ChaCha20 draws on the device); wall clock of "
                           "DeviceGSO::run incl. ld_gso_create, the state read-backs at the save points and ld_gso_destroy"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- executed work (culling) on a sample, outside the timed region -------------------------
    sample = poses[:: max(1, n_local // 256)][:256]
    det = (ldb200.PoseDetail * len(sample))()
    es = np.empty(len(sample))
    ldb200._check(lib, lib.ld_score_batch_detail(h, len(sample), np.ascontiguousarray(sample).ctypes.data,
                                                 es.ctypes.data, det, None, None))
    tested = np.mean([d.n_pairs_tested for d in det]) / (N_REC * N_LIG)
    in_cut = np.mean([d.n_in_cutoff for d in det]) / (N_REC * N_LIG)

    # ---- roofline denominators measured on this box ---------------------------------------------
    fp64_tf, fp32_tf, gather_g = ldb200.probe_peaks(local_rank)
    pairs_per_rank_step = n_local * N_REC * N_LIG
    brute_tf = 8.0 * pairs_per_rank_step / (pair_ms_step * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    sm_hz = (clk["sm_mhz"] if clk and clk.get("sm_mhz") else peaks.get("sm_max_mhz", 1965.0)) * 1e6
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    # algorithmic HBM bytes per pose on the ligand-frame path: pose row in, energy out, per-pose rotation data (128 B
    # written + read), per-group partial sums and receptor interface words (written + read), ligand interface bitmap
    # (cleared + read); the complex itself (cell lists, table rows, ligand) is L2 / shared-memory resident
    n_groups = int(case.path_info().split(" receptor groups")[0].split()[-1]) if path == "rigid" else 0
    hbm_bytes = n_local * (56 + 8 + 2 * 128 + 2 * n_groups * (8 + 4) + 2 * 4 * ((N_LIG + 7) // 8 * 8 + 31) // 32)
    counts, counts_why = load_counts()
    roofline = {"bound": "issue", "kernel": pair_kernel, "path": path, "path_info": case.path_info(),
                "achieved": None, "peak": sms * 4 * sm_hz / 1e9, "unit": "Gwarp-inst/s", "frac": None, "traffic": None}
    if counts and "1k4c_bench" in counts and path == "rigid":
        iss = issue_roofline(counts["1k4c_bench"], n_local, pair_ms_step, sm_hz, sms)
        roofline.update({"achieved": iss["achieved_gwarp_inst_per_s"], "frac": iss["frac"], "issue": iss,
                         "traffic": (counts["1k4c_bench"]["dram_read_bytes"] + counts["1k4c_bench"]["dram_write_bytes"])
                         / counts["1k4c_bench"]["poses"] * n_local / max(1, pair_launches // args.steps),
                         "counts_source": "profiles/r2_kernel_counts.json (ncu capture of this kernel on the 80,000-pose "
                                          "bench step; source hash verified); time, clock and SM count measured live"})
    else:
        roofline["counts_refused"] = counts_why or "no counters for this kernel"
    roofline.update({
        "what_bounds_it": "warp-instruction issue (4 schedulers x SMs x SM clock), co-limited by shared-memory wavefronts "
                          "(one conflict-free LDS.128 of ligand atoms + one 8-byte table gather per pair; the gather's random "
                          "bank collisions are a third of the wavefronts); FP64 pipe ~5 %, DRAM < 1 % of peak",
        "pair_kernel_ms_per_step": pair_ms_step, "pair_kernel_ms_per_step_by_rank": pair_ms_ranks,
        "pair_kernel_spread_max_over_mean": max(pair_ms_ranks) / (sum(pair_ms_ranks) / len(pair_ms_ranks)),
        "poses_by_rank": poses_ranks, "pair_kernel_launches_timed": pair_launches,
        "executed_pair_test_fraction": tested, "in_cutoff_fraction": in_cut,
        "bruteforce_equivalent": {"tflops": brute_tf, "fp64_nonfma_probe_tflops": fp64_tf, "ratio": brute_tf / fp64_tf,
                                  "note": "SURVEY.md 8(d): 8 FP64 flops x N_rec x N_lig per pose against the measured non-fused "
                                          "FP64 rate.  The kernel executes `executed_pair_test_fraction` of the pairs, in FP32 "
                                          "with proven margins, so this ratio is the pruning + reformulation factor, NOT a "
                                          "roofline fraction"},
        "gather": {"gloads_per_s": in_cut * pairs_per_rank_step / (pair_ms_step * 1e-3) / 1e9,
                   "l2_gather_probe_gloads_per_s": gather_g, "note": "table values come from shared memory on this path"},
        "hbm": {"achieved_gbs": hbm_bytes / ((ms_total / args.steps) * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                "frac": hbm_bytes / ((ms_total / args.steps) * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
        "fp32_nonfma_tflops_measured": fp32_tf,
        "share_of_step": {"prep": tr_ms / (tr_ms + pair_ms + fin_ms), "pair": pair_ms / (tr_ms + pair_ms + fin_ms),
                          "finalize": fin_ms / (tr_ms + pair_ms + fin_ms)}})

    # ---- CPU baseline (oracle port on this box's host cores, bounded sample) -------------------
    cores = host_cores()
    cpu = None
    configs = None
    if not args.no_cpu_baseline and world == 1:  # reported at N=1 only (rank 0's host cores are shared at N>1)
        cx = oracle_complex()
        rate, n_cpu, dt = cpu_rate(cx, np.ascontiguousarray(all_poses.reshape(-1, 7)[:: max(1, n_total // 8192)]),
                                   cores, args.cpu_seconds)
        e_cpu = cx.energy(poses[:8])
        rel = np.max(np.abs(e_cpu - e_dev[:8]) / np.abs(e_cpu))
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_cpu} poses of the same workload in {dt:.1f} s on {cores} threads (oracle/ld_oracle.c)",
               "pair_evals_per_s": rate * N_REC * N_LIG, "gpu_vs_oracle_max_rel_err_8_poses": float(rel), "note": CPU_NOTE}
    if world == 1 and not args.no_configs:
        configs = config_leg(local_rank, fp64_tf, sm_hz, sms, counts, cores, args.config_cpu_seconds)

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": f"synthetic swarms on the real 1k4c structures; DCparams {table_kind}",
           "config": workload_config(args, world), "pair_evals_per_s": value * N_REC * N_LIG,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_local * 7 * 8,
                   "d2h_bytes_per_step": n_local * 8, "ms_per_step": e2e_s / args.steps * 1e3,
                   "api": "ld_score_batch (C ABI, host buffers)"},
           "gpu_launches": launches_per_step * args.steps, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
           "gso_run": gso, "gso_device_run": gso_dev, "configs": configs}
    if world == 1 and not args.no_single_swarm_runs:
        out["single_swarm_runs"] = single_swarm_runs(dc_dir)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# README.md:27-147 of the reference: whole-run wall clock, 1 swarm x 200 glowworms x 100 steps, one M3 Pro core
README_M3_SECONDS = {"1k4c": 112.132, "1ppe": 4.252, "2uuy": 8.108, "1czy": 1.580, "1azp": 14.228}


def single_swarm_runs(dc_dir, repeats=5):
    """BASELINE configs 0-3 as a reference user runs them: the drop-in CLI, one swarm, 100 steps, whole-process wall
    clock (process start, CUDA context, model building and the 11 output files included), next to the README's M3 Pro
    times, with the CLI's own account of where the time went (LDB200_TIMING=1: one line on stderr).  Each configuration
    runs `repeats` times (a run is >90 % process and CUDA start-up, whose duration is the box's and varies from 0.2 to
    1.7 s between consecutive runs): `wall_s` is the median run, `wall_s_min` the fastest, `wall_s_runs` lists them all,
    `breakdown` is the median run's."""
    import shutil
    from ldb200 import host
    golden = os.path.join(ROOT, "tests", "golden")

    def run_cli(name, method, device_gso):
        g = os.path.join(golden, name)
        start = os.path.join(g, "initial_positions_0.dat")
        if not os.path.exists(start):
            start = os.path.join(g, "init", "initial_positions_0.dat")
        with tempfile.TemporaryDirectory() as tmp:
            for f in ("rec_nm.npy", "lig_nm.npy"):
                if os.path.exists(os.path.join(g, f)):
                    shutil.copy(os.path.join(g, f), os.path.join(tmp, f))
            env = dict(os.environ, LIGHTDOCK_DATA=dc_dir, LDB200_TIMING="1")
            if device_gso:
                env["LIGHTDOCK_GSO"] = "device"  # the GSO step on the GPU too
            t = time.perf_counter()
            r = subprocess.run([host.CLI_PATH, os.path.join(g, "setup.json"), start, "100", method], cwd=tmp, env=env,
                               capture_output=True, text=True)
            dt = time.perf_counter() - t
            ok = r.returncode == 0 and os.path.exists(os.path.join(tmp, "swarm_0", "gso_100.out"))
        return (dt if ok else None), r.stderr

    def median_run(name, method, device_gso):
        runs = [run_cli(name, method, device_gso) for _ in range(repeats)]
        good = sorted((r for r in runs if r[0] is not None), key=lambda r: r[0])
        if not good:
            return None, "", [None] * repeats
        return good[len(good) // 2][0], good[len(good) // 2][1], [r[0] for r in runs]

    out = {}
    run_cli("1czy", "dfire", False)  # the first process of a fresh box pays one-off costs (driver / page cache)
    for name, method in (("1czy", "dfire"), ("1ppe", "dfire"), ("2uuy", "dfire"), ("1azp", "dna"), ("1k4c", "dfire")):
        dt, stderr, all_dt = median_run(name, method, False)
        breakdown = None
        m = re.search(r"\[ldb200 timing\] (.*)", stderr)
        if m and dt is not None:
            nums = dict((k2, float(v)) for k2, v in re.findall(r"(\w+)=([0-9.]+)", m.group(1)))
            in_main = nums.get("total_in_main_ms", 0.0)
            breakdown = {"process_start_and_exit_ms": dt * 1e3 - in_main,   # exec, dynamic loading (libcudart), teardown
                         "parse_inputs_ms": nums.get("until_main_inputs_ms"),
                         "load_case_ms": nums.get("load_case_ms"),
                         "of_which_ld_create": {"cuda_context_wait_ms": nums.get("context_wait"), "complex_ms": nums.get("complex"),
                                                "receptor_groups_ms": nums.get("groups"), "cell_lists_ms": nums.get("cells")},
                         "gso_100_steps_ms": nums.get("gso_ms"), "energy_calls": int(nums.get("energy_calls", 0))}
        out[name] = {"method": method, "wall_s": dt, "wall_s_min": min((x for x in all_dt if x is not None), default=None),
                     "wall_s_runs": all_dt, "readme_m3pro_1core_wall_s": README_M3_SECONDS[name],
                     "breakdown": breakdown}
        dt, stderr, all_dt = median_run(name, method, True)
        m = re.search(r"gso_ms=([0-9.]+)", stderr)
        out[name]["device_gso"] = {"wall_s": dt, "wall_s_runs": all_dt, "gso_100_steps_ms": float(m.group(1)) if m else None}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--swarms", type=int, default=400)
    ap.add_argument("--glowworms", type=int, default=200)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--config-cpu-seconds", type=float, default=1.5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE configs 0-3 leg reported at N=1")
    ap.add_argument("--no-single-swarm-runs", action="store_true",
                    help="skip the five 100-step single-swarm CLI runs (BASELINE configs 0-3) reported at N=1")
    ap.add_argument("--gso-steps", type=int, default=100,
                    help="steps of the real GSO loop for the gso_run figures (BASELINE configs[4]: 100; 0 = skip)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
