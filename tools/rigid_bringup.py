"""GPU bring-up of the rigid-ligand DFIRE path (not a test): parity against the oracle and against the
generic path, work counters, and timing of both paths on the bench workload.  Run under gpurun."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("", "oracle", "lightdock-rust_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import ldb200  # noqa: E402
import oracle as O  # noqa: E402
from helpers import assert_parity, case, random_poses, scorer_from_oracle  # noqa: E402
from ldb200 import workload  # noqa: E402

for name in ("1ppe", "1k4c"):
    cx, pos, _ = case(name, O.DFIRE)
    sc = scorer_from_oracle(cx)
    print(name, sc.path_info())
    rng = np.random.default_rng(3)
    close = random_poses(rng, 12, cx.pose_len, centre=cx.rec.coords.mean(axis=0), spread=12.0)
    for label, poses in (("start", pos[:16]), ("close", close)):
        e_ref, d_ref = cx.energy(poses, detail=True)
        for path in (ldb200.PATH_RIGID, ldb200.PATH_GENERIC):
            sc.set_path(path)
            e, d = sc.energy_detail(poses)
            try:
                rel = assert_parity(e, d, e_ref, d_ref, cx.method)
                ok = f"parity OK rel {rel:.1e}"
            except AssertionError as ex:
                ok = "PARITY FAIL: " + str(ex).strip().splitlines()[0][:200]
            same = np.array_equal(sc.energy(poses), e)
            print(f"  {label:5s} path {sc.stats()['path']}: {ok}; plain==detail {same}; tested/in-cutoff "
                  f"{d['n_pairs_tested'].sum() / max(1, d['n_in_cutoff'].sum()):.2f}; exact-fallback "
                  f"{d['n_exact_fallback'].sum() / max(1, d['n_in_cutoff'].sum()):.4f}")

cx, pos, _ = case("1k4c", O.DFIRE)
sc = scorer_from_oracle(cx)
n_sw = int(os.environ.get("SWARMS", "100"))
poses = np.ascontiguousarray(workload.synthetic_1k4c_swarms(n_sw, 200).reshape(-1, 7))
sc.set_profiling(True)
for path in (ldb200.PATH_GENERIC, ldb200.PATH_RIGID):
    sc.set_path(path)
    sc.energy(poses[:2000])
    for rep in range(2):
        t = time.time()
        e = sc.energy(poses)
        dt = time.time() - t
        st = sc.stats()
        print(f"path {st['path']}: {len(poses)} poses wall {dt * 1e3:.1f} ms device {st['device_ms']:.1f} ms "
              f"(transform {st['transform_ms']:.2f} pair {st['pair_ms']:.2f} finalize {st['finalize_ms']:.2f}) -> "
              f"{len(poses) / st['device_ms'] * 1e3:.0f} poses/s")
    if path == ldb200.PATH_GENERIC:
        e_gen = e
print("rigid vs generic max rel diff", np.max(np.abs(e - e_gen) / np.abs(e_gen)))
