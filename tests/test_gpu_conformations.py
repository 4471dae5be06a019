"""GPU: bin/lightdock-rust-conformations (pose output) writes the complexes of a gso_<step>.out with the
library's own pose transform.  Checked against the oracle's transformed coordinates at PDB precision."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle as O
from helpers import GOLDEN, case

pytestmark = pytest.mark.gpu


def _atoms(path):
    rec, cur = [], []
    parts = []
    for line in open(path):
        if line.startswith(("ATOM", "HETATM")):
            cur.append((line[12:16], line[17:20], line[21], line[22:27], float(line[30:38]), float(line[38:46]), float(line[46:54])))
        elif line.startswith("TER"):
            parts.append(cur)
            cur = []
    assert not cur, "every partner ends with TER"
    return parts


@pytest.mark.parametrize("name,method", [("1azp", "dna"), ("1k4c", "dfire")])
def test_conformations_match_the_oracle_transform(name, method, tmp_path, monkeypatch):
    from ldb200 import host, workload
    g = os.path.join(GOLDEN, name)
    cx, _, _ = case(name, O.DNA if method == "dna" else O.DFIRE)
    env = dict(os.environ)
    if method == "dfire":
        dc_dir, _ = workload.ensure_dcparams_dir(str(tmp_path))
        env["LIGHTDOCK_DATA"] = dc_dir
    for f in ("rec_nm.npy", "lig_nm.npy"):
        if os.path.exists(os.path.join(g, f)):
            shutil.copy(os.path.join(g, f), tmp_path / f)
    os.makedirs(tmp_path / "swarm_0")
    shutil.copy(os.path.join(g, "swarm_0", "gso_100.out"), tmp_path / "swarm_0" / "gso_100.out")
    exe = os.path.join(os.path.dirname(host.CLI_PATH), "lightdock-rust-conformations")
    ids = [0, 7, 199]
    r = subprocess.run([exe, os.path.join(g, "setup.json"), "swarm_0/gso_100.out", method] + [str(i) for i in ids],
                       cwd=tmp_path, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == "Wrote 3 structures next to swarm_0/gso_100.out"
    poses = O.parse_gso_out(os.path.join(g, "swarm_0", "gso_100.out"))[0]
    _, d = cx.energy(poses[ids], detail=True)
    for k, i in enumerate(ids):
        rec, lig = _atoms(tmp_path / "swarm_0" / f"lightdock_{i}.pdb")
        assert len(rec) == cx.rec.n and len(lig) == cx.lig.n
        got_r = np.array([a[4:] for a in rec]); got_l = np.array([a[4:] for a in lig])
        assert np.abs(got_r - d["coords_rec"][k]).max() <= 5.01e-4
        assert np.abs(got_l - d["coords_lig"][k]).max() <= 5.01e-4
    # the records are the input's (names, residues, chains), in model order
    import json
    setup = json.load(open(os.path.join(g, "setup.json")))
    for part, key in ((rec, "receptor_pdb"), (lig, "ligand_pdb")):
        src = [(l[12:16], l[17:20], l[21], l[22:27]) for l in open(os.path.join(g, "lightdock_" + setup[key]))
               if l.startswith(("ATOM", "HETATM"))]
        assert [a[:4] for a in part] == src
    r = subprocess.run([exe, os.path.join(g, "setup.json"), "swarm_0/gso_100.out", method, "200"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 1 and "no glowworm 200" in r.stderr
