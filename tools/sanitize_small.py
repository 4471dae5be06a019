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
    if name == "1czy":  # FLEX: the first call was the brute-force route, this one goes through the rebuilt slack lists
        print("   flex lists", float(sc.energy(poses)[0]), sc.path_info()[-60:])

# device-resident GSO (gso_step_kernel + live-row-count scoring): 2 swarms x 40 glowworms, 6 steps, DNA with ANM
import ctypes as C
cx, pos, _ = case("1azp", O.DNA)
sc = scorer_from_oracle(cx)
lib = ldb200.load_library()
start = np.ascontiguousarray(np.stack([pos[:40], pos[40:80]]), np.float64)
seeds = np.array([324324, 7], np.uint64)
g = C.c_void_p()
assert lib.ld_gso_create(sc.h, 2, 40, start.ctypes.data, seeds.ctypes.data, C.byref(g)) == 0, lib.ld_last_error()
assert lib.ld_gso_run(g, 6) == 0, lib.ld_last_error()
out = np.empty_like(start)
assert lib.ld_gso_state(g, out.ctypes.data, None, None, None, None, None) == 0
print("device gso: steps", lib.ld_gso_steps(g), "energy calls", lib.ld_gso_energy_calls(g), "moved", float(np.abs(out - start).max()) > 0)
lib.ld_gso_destroy(g)
