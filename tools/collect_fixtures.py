#!/usr/bin/env python3
"""Copy the reference's DATA fixtures (never sources) into tests/golden/.

The GPU box has no /root/reference, so every input structure, start-position file, ANM file and
golden `gso_*.out` trajectory that the parity tests need is committed under tests/golden/.
Run once in the build container:   python tools/collect_fixtures.py

What is collected (all are inputs/outputs of the reference, no code):
  tests/{1azp,2oob}/*.pdb                     known-answer structures (src/dna.rs:544-552, src/dfire.rs:388-396)
  example/<case>/setup.json, lightdock_*.pdb, initial_positions_0.dat, rec_nm.npy, lig_nm.npy
  example/<case>/swarm_0/gso_*.out            golden trajectories
"""
import os, shutil, sys, glob

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "..", "tests", "golden")

CASES = {
    "1azp": ["setup.json", "initial_positions_0.dat", "lightdock_protein.pdb", "lightdock_dna.pdb",
             "rec_nm.npy", "lig_nm.npy"],
    "1k4c": ["setup.json", "initial_positions_0.dat", "lightdock_receptor_membrane.pdb",
             "lightdock_ligand.pdb"],
    "1ppe": ["setup.json", "initial_positions_0.dat", "lightdock_1ppe_e.pdb", "lightdock_1ppe_i.pdb"],
    "2uuy": ["setup.json", "initial_positions_0.dat", "lightdock_2UUY_rec.pdb", "lightdock_2UUY_lig.pdb",
             "rec_nm.npy", "lig_nm.npy"],
    "1czy": ["setup.json", "init/initial_positions_0.dat", "lightdock_1czy_protein.pdb",
             "lightdock_1czy_peptide.pdb", "rec_nm.npy", "lig_nm.npy"],
    "ab_icode": ["setup.json", "init/initial_positions_0.dat", "lightdock_receptor.pdb",
                 "lightdock_ligand.pdb", "rec_nm.npy", "lig_nm.npy"],
}


def cp(src, dst):
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copyfile(src, dst)


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; fixtures are already committed")
    for t in ("1azp", "2oob"):
        for f in glob.glob(f"{REF}/tests/{t}/*.pdb"):
            cp(f, os.path.join(DST, "unit", t, os.path.basename(f)))
    for case, files in CASES.items():
        for f in files:
            cp(f"{REF}/example/{case}/{f}", os.path.join(DST, case, f))
        for f in glob.glob(f"{REF}/example/{case}/swarm_0/gso_*.out"):
            cp(f, os.path.join(DST, case, "swarm_0", os.path.basename(f)))
    print("fixtures written to", os.path.normpath(DST))


if __name__ == "__main__":
    main()
