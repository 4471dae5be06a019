import os, sys, time
import numpy as np
ROOT = os.getcwd()
for p in ("", "oracle", "lightdock-rust_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import oracle as O
from helpers import case, scorer_from_oracle
from ldb200 import workload
cx, pos, _ = case("1k4c", O.DFIRE)
sc = scorer_from_oracle(cx)
sc.set_profiling(True)
for nsw in (50, 400):
    poses = np.ascontiguousarray(workload.synthetic_1k4c_swarms(400, 200)[::400 // nsw].reshape(-1, 7))
    sc.energy(poses)
    ts = []
    for _ in range(5):
        sc.energy(poses); ts.append(sc.stats()["pair_ms"])
    print(os.environ.get("LDB200_UNITS_PER_SM"), len(poses), "poses: pair ms", np.round(ts, 3), "->", len(poses) / np.median(ts) * 1e3 / 1e6, "M poses/s")
