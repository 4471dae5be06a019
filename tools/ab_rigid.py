"""A/B of one build of the CUDA library (LDB200_LIB=<variant .so>) on the headline workload: parity of the rigid
kernel against the oracle (16 start poses + 12 close poses, integer outputs bit-exact) and against the generic kernel
(2,000 bench poses, energies), then the pair-kernel time on 10,000 and 80,000 bench poses and on 20,000 poses of the
FLEX configurations.  Run under gpurun."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("", "oracle", "lightdock-rust_b200", "tests", "tools"):
    sys.path.insert(0, os.path.join(ROOT, p))
import ldb200  # noqa: E402
import oracle as O  # noqa: E402
from helpers import assert_parity, case, random_poses, scorer_from_oracle  # noqa: E402
from ldb200 import workload  # noqa: E402
from count_target import config_poses  # noqa: E402

tag = os.path.basename(os.environ.get("LDB200_LIB", "default"))
cx, pos, _ = case("1k4c", O.DFIRE)
sc = scorer_from_oracle(cx)
rng = np.random.default_rng(3)
close = random_poses(rng, 12, cx.pose_len, centre=cx.rec.coords.mean(axis=0), spread=12.0)
for label, poses in (("start", pos[:16]), ("close", close)):
    e_ref, d_ref = cx.energy(poses, detail=True)
    e, d = sc.energy_detail(poses)
    rel = assert_parity(e, d, e_ref, d_ref, cx.method)
    assert np.array_equal(sc.energy(poses), e)
    print(f"{tag}: {label} parity OK rel {rel:.1e} exact-fallback {d['n_exact_fallback'].sum() / max(1, d['n_in_cutoff'].sum()):.5f}")
bench = np.ascontiguousarray(workload.synthetic_1k4c_swarms(400, 200).reshape(-1, 7))
sub = bench[::40]
e_r = sc.energy(sub)
sc.set_path(ldb200.PATH_GENERIC)
e_g = sc.energy(sub)
sc.set_path(ldb200.PATH_RIGID)
print(f"{tag}: rigid vs generic on {len(sub)} bench poses: max rel diff {np.max(np.abs(e_r - e_g) / np.abs(e_g)):.2e}")
sc.set_profiling(True)
for n in (10000, 80000):
    poses = np.ascontiguousarray(bench[::len(bench) // n])
    sc.energy(poses)
    ts = []
    for _ in range(5):
        sc.energy(poses)
        ts.append(sc.stats()["pair_ms"])
    print(f"{tag}: 1k4c {len(poses)} poses: pair ms {np.round(ts, 3)} -> {len(poses) / np.median(ts) * 1e3 / 1e6:.3f} M poses/s")
# what one rank of an 8-GPU run scores: 50 whole swarms under the cost-aware map (contiguous blocks of 200 similar poses)
all_sw = workload.synthetic_1k4c_swarms(400, 200)
rec_xyz = workload.read_pdb_coords(os.path.join(workload.GOLDEN_1K4C, "lightdock_receptor_membrane.pdb"))
lig_xyz = workload.read_pdb_coords(os.path.join(workload.GOLDEN_1K4C, "lightdock_ligand.pdb"))
for world, rank in ((8, 0), (8, 5), (4, 1)):
    mine = workload.shard_swarms_cost_aware(all_sw, rec_xyz, lig_xyz, rank, world)
    poses = np.ascontiguousarray(all_sw[mine].reshape(-1, 7))
    sc.energy(poses)
    ts = []
    for _ in range(5):
        sc.energy(poses)
        ts.append(sc.stats()["pair_ms"])
    print(f"{tag}: 1k4c rank {rank} of {world} ({len(poses)} poses, whole swarms): pair ms {np.round(ts, 3)} -> x{world} = {np.median(ts) * world:.3f} ms")
sc.close()
for name in ("1ppe", "2uuy", "1czy"):
    cxc, poses = config_poses(name)
    s2 = scorer_from_oracle(cxc)
    s2.set_profiling(True)
    s2.energy(poses); s2.energy(poses)
    ts = []
    for _ in range(5):
        s2.energy(poses)
        ts.append(s2.stats()["pair_ms"])
    print(f"{tag}: {name} {len(poses)} poses: pair ms {np.round(ts, 3)} -> {len(poses) / np.median(ts) * 1e3 / 1e6:.3f} M poses/s")
    s2.close()
