#!/usr/bin/env python3
"""Generate the swarm centres of the synthetic 1k4c-shaped benchmark (BASELINE.json configs[4]).

400 centres on the extracellular side of the 1k4c membrane receptor, accepted when random poses drawn
around them (translation ~ N(centre, 4.4 A) clipped to a 10 A sphere, uniform unit quaternions — the
statistics of example/1k4c/initial_positions_0.dat) have 0.4-2 % of their atom pairs inside DFIRE's
15 A cut-off, the range measured on the real swarm.  Output: one "x y z" row per swarm in
lightdock-rust_b200/data/synthetic_1k4c_swarm_centres.tsv.  Seeded; rerunning reproduces the file.
"""
import os
import sys

import numpy as np
from scipy.spatial import cKDTree

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "lightdock-rust_b200"))
from ldb200.workload import read_pdb_coords, swarm_poses, rotate_points  # noqa: E402

GOLD = os.path.join(HERE, "..", "tests", "golden", "1k4c")
N_SWARMS = 400


def main():
    rec = read_pdb_coords(os.path.join(GOLD, "lightdock_receptor_membrane.pdb"))
    lig = read_pdb_coords(os.path.join(GOLD, "lightdock_ligand.pdb"))
    tree = cKDTree(rec)
    rng = np.random.default_rng(324324)
    centres, fracs = [], []
    while len(centres) < N_SWARMS:
        phi = rng.uniform(0, 2 * np.pi)
        r = rng.uniform(0, 75.0)
        z = rng.uniform(24.0, 75.0)
        c = np.array([r * np.cos(phi), r * np.sin(phi), z])
        poses = swarm_poses(rng, c, 6)
        f = []
        for p in poses:
            xyz = rotate_points(p[3:7], lig) + p[:3]
            f.append(tree.count_neighbors(cKDTree(xyz), 15.0) / (len(rec) * len(lig)))
        f = np.array(f)
        if 0.004 <= f.mean() <= 0.02 and f.min() >= 0.002:
            centres.append(c)
            fracs.append(f.mean())
    out = os.path.join(HERE, "..", "lightdock-rust_b200", "data", "synthetic_1k4c_swarm_centres.tsv")
    with open(out, "w") as fh:
        fh.write("# swarm centres (x y z, Angstrom) of the synthetic 1k4c-shaped benchmark; tools/make_synthetic_swarms.py\n")
        for c in centres:
            fh.write(f"{c[0]:.3f}\t{c[1]:.3f}\t{c[2]:.3f}\n")
    print(f"{len(centres)} centres, mean in-cutoff fraction {np.mean(fracs):.4f} (min {np.min(fracs):.4f}, max {np.max(fracs):.4f})")


if __name__ == "__main__":
    main()
