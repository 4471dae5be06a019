"""Per-configuration timings (run under gpurun): single-swarm call latency and large-batch throughput of
every BASELINE config through the C ABI with host buffers."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("", "oracle", "lightdock-rust_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import oracle as O  # noqa: E402
from helpers import case, random_poses, scorer_from_oracle  # noqa: E402

rows = []
for name, method in (("1czy", O.DFIRE), ("1ppe", O.DFIRE), ("2uuy", O.DFIRE), ("1azp", O.DNA), ("1k4c", O.DFIRE)):
    cx, pos, _ = case(name, method)
    sc = scorer_from_oracle(cx)
    swarm = pos[:200]
    sc.energy(swarm)
    t = time.perf_counter()
    for _ in range(50):
        sc.energy(swarm)
    lat = (time.perf_counter() - t) / 50
    rng = np.random.default_rng(1)
    big = np.tile(pos, (100, 1))
    big[:, :3] += rng.normal(0, 1.0, size=(len(big), 3))
    sc.energy(big)
    t = time.perf_counter()
    sc.energy(big)
    dt = time.perf_counter() - t
    st = sc.stats()
    print(f"{name:5s} {('DFIRE' if method == O.DFIRE else 'DNA'):5s} rec {cx.rec.n:5d} lig {cx.lig.n:5d} modes {cx.rec.n_modes if cx.use_anm else 0}/"
          f"{cx.lig.n_modes if cx.use_anm else 0}  path {st['path']}  200-pose call {lat * 1e6:7.1f} us   {len(big)} poses: "
          f"{len(big) / dt / 1e6:6.3f} M poses/s  {len(big) / dt * cx.rec.n * cx.lig.n / 1e12:7.3f} T pair-evals/s", flush=True)
