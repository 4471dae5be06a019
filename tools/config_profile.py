"""Per-configuration kernel breakdown (run under gpurun): per-kernel device time of a 20,000-pose batch, executed-pair
fraction and in-cut-off fraction, for every BASELINE config and both DFIRE paths where they apply."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("", "oracle", "lightdock-rust_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import ldb200  # noqa: E402
import oracle as O  # noqa: E402
from helpers import case, scorer_from_oracle  # noqa: E402

for name, method in (("1czy", O.DFIRE), ("1ppe", O.DFIRE), ("2uuy", O.DFIRE), ("ab_icode", O.DFIRE), ("1azp", O.DNA), ("1k4c", O.DFIRE)):
    cx, pos, _ = case(name, method)
    sc = scorer_from_oracle(cx)
    rng = np.random.default_rng(1)
    big = np.tile(pos, (100, 1))
    big[:, :3] += rng.normal(0, 1.0, size=(len(big), 3))
    info = sc.path_info()
    paths = [ldb200.PATH_GENERIC] + ([ldb200.PATH_RIGID] if info.startswith("rigid path on") else [])
    for path in paths:
        sc.set_path(path)
        sc.set_profiling(True)
        sc.energy(big)
        t = time.perf_counter(); sc.energy(big); dt = time.perf_counter() - t
        st = sc.stats()
        _, d = sc.energy_detail(big[:1000])
        n = cx.rec.n * cx.lig.n
        print(f"{name:8s} {('DFIRE' if method == O.DFIRE else 'DNA'):5s} rec {cx.rec.n:5d} lig {cx.lig.n:5d} path {st['path']} "
              f"{len(big)} poses {dt*1e3:7.2f} ms ({len(big)/dt/1e6:6.3f} M poses/s)  transform {st['transform_ms']:6.2f} pair {st['pair_ms']:6.2f} "
              f"finalize {st['finalize_ms']:5.2f} ms  tested {d['n_pairs_tested'].mean()/n:.3f} in-cutoff {d['n_in_cutoff'].mean()/n:.3f} "
              f"exact {d['n_exact_fallback'].sum()/max(1,d['n_in_cutoff'].sum()):.5f}", flush=True)
    print("         ", info[:150])
