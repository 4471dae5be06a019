"""Large randomized cross-check (run under gpurun): the two DFIRE kernels classify every atom pair along independent
routes (ligand frame + cell lists vs lab frame + sphere culling); their bin histograms, in-cut-off counts and
interface flags must be identical on every pose.  Prints the number of poses / in-cut-off pairs compared."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("", "oracle", "lightdock-rust_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import ldb200  # noqa: E402
import oracle as O  # noqa: E402
from helpers import case, random_poses, scorer_from_oracle  # noqa: E402
from ldb200 import workload  # noqa: E402

N = int(os.environ.get("POSES", "40000"))
total_pairs = 0
total_fallback = [0, 0]
worst = 0.0
t0 = time.time()
for name in ("1k4c", "1ppe"):
    cx, pos, _ = case(name, O.DFIRE)
    sc = scorer_from_oracle(cx)
    rng = np.random.default_rng(2024)
    sets = [random_poses(rng, N // 2, 7, centre=cx.rec.coords.mean(axis=0), spread=14.0)]
    if name == "1k4c":
        sets.append(workload.synthetic_1k4c_swarms(400, 200).reshape(-1, 7)[rng.permutation(80000)[:N // 2]])
    else:
        sets.append(np.tile(pos, (N // 2 // len(pos) + 1, 1))[:N // 2] + np.concatenate(
            [rng.normal(0, 1.5, size=(N // 2, 3)), np.zeros((N // 2, 4))], axis=1))
    for poses in sets:
        for lo in range(0, len(poses), 4000):
            chunk = np.ascontiguousarray(poses[lo:lo + 4000])
            sc.set_path(ldb200.PATH_RIGID)
            e_r, d_r = sc.energy_detail(chunk)
            sc.set_path(ldb200.PATH_GENERIC)
            e_g, d_g = sc.energy_detail(chunk)
            for k in ("n_in_cutoff", "n_interface_pairs", "bin_hist", "rec_rst_hit", "lig_rst_hit", "membrane_hit",
                      "iface_rec", "iface_lig"):
                if not np.array_equal(d_r[k], d_g[k]):
                    bad = np.where((d_r[k] != d_g[k]).reshape(len(chunk), -1).any(axis=1))[0]
                    raise SystemExit(f"MISMATCH in {k} for {name}, poses {lo + bad[:5]}")
            # the kernels add the same table values in different orders: compare against the size of the SUM (the
            # score (sum*0.0157 - 4.7) * -1 can cancel to ~0, where a relative test on the score is meaningless)
            scale = np.maximum(np.abs(d_g["raw_sum"]) * 0.0157, 1.0)
            err = np.abs(e_r - e_g) / scale
            worst = max(worst, float(err.max()))
            assert err.max() < 1e-10, err.max()
            total_pairs += int(d_r["n_in_cutoff"].sum())
            total_fallback[0] += int(d_r["n_exact_fallback"].sum())
            total_fallback[1] += int(d_g["n_exact_fallback"].sum())
    print(f"{name}: {2 * (N // 2)} poses identical so far; cumulative in-cut-off pairs {total_pairs:,}", flush=True)
print(f"OK: {total_pairs:,} in-cut-off pair classifications identical between the rigid and the generic kernel; "
      f"exact-path pairs rigid {total_fallback[0]:,} / generic {total_fallback[1]:,}; largest energy difference "
      f"{worst:.1e} of the summed magnitude; {time.time() - t0:.0f} s")

# ---- second leg: the rigid kernel against the CPU oracle (energies only, all host cores) --------------------------
ncores = len(os.sched_getaffinity(0))
for name, n in (("1ppe", 60000), ("1k4c", 12000)):
    cx, pos, _ = case(name, O.DFIRE)
    sc = scorer_from_oracle(cx)
    sc.set_path(ldb200.PATH_RIGID)
    rng = np.random.default_rng(77)
    poses = np.vstack([random_poses(rng, n // 2, 7, centre=cx.rec.coords.mean(axis=0), spread=14.0),
                       np.tile(pos, (n // 2 // len(pos) + 1, 1))[:n // 2] + np.concatenate(
                           [rng.normal(0, 1.5, size=(n // 2, 3)), np.zeros((n // 2, 4))], axis=1)])
    t = time.time()
    e_ref = cx.energy_mt(poses, ncores)
    dt = time.time() - t
    e_gpu = sc.energy(poses)
    # tolerance of the north star: 1e-6 relative; scores that cancel to ~0 are judged against the sum's size
    err = np.abs(e_gpu - e_ref) / np.maximum(np.abs(e_ref), 1.0)
    assert err.max() < 1e-9, err.max()
    print(f"{name}: {n} poses, rigid kernel vs oracle ({ncores} threads, {dt:.0f} s): max |dE|/max(|E|,1) = {err.max():.1e}", flush=True)

# ---- third leg (round 2): the FLEX instance (ligand with ANM modes) against the generic kernel and the oracle ----------
NF = int(os.environ.get("FLEX_POSES", str(N // 4)))
flex_pairs = 0
for name in ("2uuy", "1czy", "ab_icode"):
    cx, pos, _ = case(name, O.DFIRE)
    sc = scorer_from_oracle(cx)
    assert sc.path_info().startswith("rigid path on (flexible ligand"), sc.path_info()
    rng = np.random.default_rng(99)
    poses = np.tile(pos, (NF // len(pos) + 1, 1))[:NF].copy()
    poses[:, :3] += rng.normal(0, 2.0, size=(NF, 3))
    q = poses[:, 3:7] + rng.normal(0, 0.15, size=(NF, 4))
    poses[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    poses[:, 7:] *= rng.uniform(0.3, 1.6, size=(NF, 1))  # ANM extents from a third of to 1.6x the shipped ones
    worst_f = 0.0
    for lo in range(0, NF, 4000):
        chunk = np.ascontiguousarray(poses[lo:lo + 4000])
        sc.set_path(ldb200.PATH_RIGID)
        e_f, d_f = sc.energy_detail(chunk)
        sc.set_path(ldb200.PATH_GENERIC)
        e_g, d_g = sc.energy_detail(chunk)
        for k in ("n_in_cutoff", "n_interface_pairs", "bin_hist", "rec_rst_hit", "lig_rst_hit", "membrane_hit",
                  "iface_rec", "iface_lig"):
            if not np.array_equal(d_f[k], d_g[k]):
                bad = np.where((d_f[k] != d_g[k]).reshape(len(chunk), -1).any(axis=1))[0]
                raise SystemExit(f"MISMATCH in {k} for {name} (FLEX), poses {lo + bad[:5]}")
        scale = np.maximum(np.abs(d_g["raw_sum"]) * 0.0157, 1.0)
        worst_f = max(worst_f, float((np.abs(e_f - e_g) / scale).max()))
        flex_pairs += int(d_f["n_in_cutoff"].sum())
    assert worst_f < 1e-9, worst_f
    n_or = min(NF, 20000)
    e_ref = cx.energy_mt(poses[:n_or], ncores)
    sc.set_path(ldb200.PATH_RIGID)
    err = np.abs(sc.energy(poses[:n_or]) - e_ref) / np.maximum(np.abs(e_ref), 1.0)
    assert err.max() < 1e-9, err.max()
    print(f"{name}: FLEX vs generic on {NF} poses identical (cumulative in-cut-off pairs {flex_pairs:,}, largest energy "
          f"difference {worst_f:.1e} of the summed magnitude); FLEX vs oracle on {n_or} poses: max |dE|/max(|E|,1) = "
          f"{err.max():.1e}; {sc.path_info().split('tile slack')[-1].strip()}", flush=True)
