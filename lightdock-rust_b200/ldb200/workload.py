"""Synthetic workloads for bench.py / smoke(): the 1k4c-shaped swarm set of BASELINE.json configs[4]
and the seeded stand-in for the missing DFIRE table.  Pure input generation (numpy); no scoring."""
import os

import numpy as np

from . import PKG_DIR

DATA_DIR = os.path.join(PKG_DIR, "data")
GOLDEN_1K4C = os.path.join(os.path.dirname(PKG_DIR), "tests", "golden", "1k4c")


def synthetic_dcparams(seed=20240324):
    """Seeded stand-in for data/DCparams (absent from the reference checkout): 169*169*20 values,
    10.0 at the shortest bins, O(1) values after, zero rows/columns for the unused type 168.
    Rounded to 9 decimals so the text file round-trips exactly."""
    rng = np.random.Generator(np.random.PCG64(seed))
    t = rng.uniform(-2.0, 2.0, size=(169, 169, 20))
    t[:, :, 0:2] = 10.0
    t[:, :, 2] = rng.uniform(0.5, 6.0, size=(169, 169))
    t[:, :, 19] *= 0.05
    t[168, :, :] = 0.0
    t[:, 168, :] = 0.0
    return np.round(t.reshape(-1), 9)


def write_dcparams(path, table):
    with open(path, "w") as f:
        f.write("\n".join(f"{v:.9f}" for v in table))
        f.write("\n")


def ensure_dcparams_dir(tmp_root):
    """Directory holding a DCparams file: $LIGHTDOCK_DATA if it has a real one, else a synthetic one
    written under tmp_root.  Returns (dir, kind)."""
    d = os.environ.get("LIGHTDOCK_DATA")
    if d and os.path.exists(os.path.join(d, "DCparams")):
        return d, "real"
    d = os.path.join(tmp_root, "ld_b200_synthetic_dcparams")
    os.makedirs(d, exist_ok=True)
    p = os.path.join(d, "DCparams")
    if not os.path.exists(p):
        write_dcparams(p + ".tmp%d" % os.getpid(), synthetic_dcparams())
        os.replace(p + ".tmp%d" % os.getpid(), p)
    return d, "synthetic"


def read_pdb_coords(path):
    xyz = []
    with open(path) as f:
        for line in f:
            if line.startswith(("ATOM  ", "HETATM")):
                xyz.append((float(line[30:38]), float(line[38:46]), float(line[46:54])))
    return np.array(xyz)


def rotate_points(q, pts):
    """Rigid rotation by quaternion (w,x,y,z) — numpy helper for workload generation only."""
    w, x, y, z = q / np.linalg.norm(q)
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return pts @ R.T


def swarm_poses(rng, centre, n):
    """n start poses of one swarm: translation ~ N(centre, 4.4 A per axis) clipped to a 10 A sphere,
    unit quaternions uniform on S^3 (statistics of example/1k4c/initial_positions_0.dat)."""
    d = rng.normal(0.0, 4.4, size=(n, 3))
    r = np.linalg.norm(d, axis=1, keepdims=True)
    d *= np.minimum(1.0, 10.0 / np.maximum(r, 1e-12))
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.hstack([centre + d, q])


def synthetic_1k4c_swarms(n_swarms=400, n_glowworms=200, seed=324324):
    """[n_swarms][n_glowworms][7] pose rows; swarm s uses centre s of the committed centre list and an
    RNG seeded with seed + s, so any rank can generate exactly its own swarms."""
    centres = np.loadtxt(os.path.join(DATA_DIR, "synthetic_1k4c_swarm_centres.tsv"), comments="#").reshape(-1, 3)
    out = np.empty((n_swarms, n_glowworms, 7))
    for s in range(n_swarms):
        rng = np.random.default_rng(seed + s)
        out[s] = swarm_poses(rng, centres[s % len(centres)], n_glowworms)
    return out


def shard_swarms(n_swarms, rank, world):
    """Swarm s -> GPU s mod G (the round-1 map): the swarm ids owned by `rank`."""
    return list(range(rank, n_swarms, world))


def shard_swarms_cost_aware(poses, rec_xyz, lig_xyz, rank, world):
    """The swarm ids owned by `rank` under the cost-aware deterministic map of host/sharding.hpp (what
    lightdock-rust-multi uses): expected in-reach atom pairs at each swarm's centre, longest-processing-time first.
    poses [n_swarms][n_glowworms][pose_len].  Every rank computes the same map; no communication."""
    from . import host
    if world == 1:
        return list(range(len(poses)))
    centres = np.asarray(poses)[:, :, :3].mean(axis=1)
    gpu, _ = host.shard_swarms(rec_xyz, lig_xyz, centres, world)
    return [s for s in range(len(poses)) if gpu[s] == rank]
