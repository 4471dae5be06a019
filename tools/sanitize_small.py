"""Small invocations of every kernel for compute-sanitizer (run under gpurun)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("", "oracle", "lightdock-rust_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import ldb200
import oracle as O
from helpers import case, scorer_from_oracle, random_poses
for name, method in (("1ppe", O.DFIRE), ("1k4c", O.DFIRE), ("1czy", O.DFIRE), ("1azp", O.DNA)):
    cx, pos, _ = case(name, method)
    sc = scorer_from_oracle(cx)
    rng = np.random.default_rng(2)
    poses = np.vstack([pos[:40], random_poses(rng, 24, cx.pose_len, centre=cx.rec.coords.mean(axis=0), spread=10.0)])
    e = sc.energy(poses)
    e2, d = sc.energy_detail(poses[:16])
    print(name, "path", sc.stats()["path"], "ok", np.isfinite(e).all(), float(e[0]))
    if name in ("1ppe", "1k4c"):
        sc.set_path(ldb200.PATH_GENERIC)
        print("   generic", float(sc.energy(poses)[0]))
