"""CPU ORACLE, Python side (test infrastructure, NOT product code).

Restates the *setup* half of the reference's scoring path — PDB iteration order, DFIRE atom typing,
DNA/pyDock AMBER parameterisation, restraint and membrane indexing — and binds the C restatement
of `energy()` / the GSO loop in oracle/ld_oracle.c through ctypes.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module.  The product host layer (lightdock-rust_b200/host) has its own, independently
written, implementation of everything in here.

Reference citations are relative to /root/reference.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "..", "lightdock-rust_b200", "data")
LIB_PATH = os.path.join(HERE, "libld_oracle.so")

DFIRE, DNA, PYDOCK = 0, 1, 2


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def _lib():
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    lib.oracle_rng_f64.restype = C.c_double
    return lib


# ------------------------------------------------------------------------------------------------
# parameter tables (data extracted from the reference by tools/extract_forcefield_tables.py)
def _tsv(name, ncol=2):
    rows = {}
    with open(os.path.join(DATA, name)) as f:
        for line in f:
            if line.startswith("#") or not line.strip():
                continue
            p = line.rstrip("\n").split("\t")
            if ncol == 3:
                rows[(p[0], p[1])] = p[2]
            else:
                rows[p[0]] = p[1]
    return rows


_TABLES = {}


def tables():
    if not _TABLES:
        _TABLES["dfire"] = {k: int(v) for k, v in _tsv("dfire_atom_types.tsv", 3).items()}
        _TABLES["amber"] = _tsv("amber_types.tsv")
        _TABLES["amber_pydock"] = dict(_TABLES["amber"], **_tsv("amber_types_pydock_extra.tsv"))
        _TABLES["ele"] = {k: float(v) for k, v in _tsv("ele_charges.tsv").items()}
        _TABLES["ele_pydock"] = dict(_TABLES["ele"],
                                     **{k: float(v) for k, v in _tsv("ele_charges_pydock_extra.tsv").items()})
        _TABLES["nt_ele"] = {k: float(v) for k, v in _tsv("nt_ele_charges.tsv").items()}
        _TABLES["vdw_e"] = {k: float(v) for k, v in _tsv("vdw_energy.tsv").items()}
        _TABLES["vdw_r"] = {k: float(v) for k, v in _tsv("vdw_radius.tsv").items()}
    return _TABLES


# ------------------------------------------------------------------------------------------------
# PDB reading.  The reference uses pdbtbx 0.11 (third-party, not vendored): ATOM/HETATM records are
# grouped chain -> residue(serial, insertion code) -> atom and iterated in that order
# (src/dfire.rs:133-186).  For contiguous files (all the fixtures) this is file order.
class Atom:
    __slots__ = ("name", "resname", "chain", "resseq", "icode", "x", "y", "z")


def read_pdb(path):
    """Atoms in the order `for chain in pdb.chains() { for residue in chain.residues() { for atom in residue.atoms()`
    visits them (src/dfire.rs:128-133) with pdbtbx 0.11's containers, restated from its published data model (the
    crate is not vendored under /root/reference; Cargo.toml:13 pins "0.11"): first model only; a chain per chain id in
    order of first appearance (a later record with an earlier chain's id joins that chain); inside a chain a residue
    per (serial number, insertion code) in order of first appearance; inside a residue a conformer per (residue name,
    alternative location) in order of first appearance -- atoms WITHOUT an alternative location form their own
    conformer --, and Residue::atoms() walks the conformers in that order.  Pinned only as far as the reference's
    fixtures go: every BASELINE structure is contiguous and alt-loc free except four residues of
    example/2uuy/lightdock_2UUY_lig.pdb; behaviour on other layouts is "parity unpinned" (DESIGN.md)."""
    chains = {}  # chain id -> {(resseq, icode) -> {(resname, altloc) -> [atoms]}} ; dicts keep insertion order
    with open(path) as f:
        for line in f:
            rec = line[:6]
            if rec == "ENDMDL":
                break  # first model only
            if rec not in ("ATOM  ", "HETATM"):
                continue
            a = Atom()
            a.name = line[12:16].strip()
            alt = line[16:17].strip()
            a.resname = line[17:20].strip()
            a.chain = line[21]
            a.resseq = int(line[22:26])
            a.icode = line[26].strip()
            a.x, a.y, a.z = float(line[30:38]), float(line[38:46]), float(line[46:54])
            chains.setdefault(a.chain, {}).setdefault((a.resseq, a.icode), {}).setdefault((a.resname, alt), []).append(a)
    atoms = []
    for ch in chains.values():
        for res in ch.values():
            names = {k[0] for k in res}
            if len(names) > 1:  # Residue::name() is None when conformers disagree: the reference panics (src/dfire.rs:134-137)
                raise ValueError("PDB Parsing Error: Residue name error")
            for conf in res.values():
                atoms.extend(conf)
    return atoms


def res_id(a):
    # "{chain}.{resname}.{resseq}" + insertion code (src/dfire.rs:138-141)
    return f"{a.chain}.{a.resname}.{a.resseq}{a.icode}"


class Molecule:
    """Numeric model of one partner (DFIREDockingModel / DNADockingModel)."""

    def __init__(self, atoms, method, active_restraints=(), modes=None, n_modes=0):
        t = tables()
        n = len(atoms)
        self.n = n
        self.coords = np.array([[a.x, a.y, a.z] for a in atoms], dtype=np.float64).reshape(n, 3)
        self.membrane = np.array([i for i, a in enumerate(atoms) if a.resname + a.name == "MMBBJ"], dtype=np.int32)
        # active restraints: only residues that exist get a key (src/dfire.rs:151-162)
        groups = {}
        active = set(active_restraints)
        for i, a in enumerate(atoms):
            rid = res_id(a)
            if rid in active:
                groups.setdefault(rid, []).append(i)
        self.rst_names = list(groups.keys())
        off = [0]
        idx = []
        for g in groups.values():
            idx.extend(g)
            off.append(len(idx))
        self.rst_offsets = np.array(off, dtype=np.int32)
        self.rst_atoms = np.array(idx, dtype=np.int32)
        self.dfire_type = None
        self.ele = self.vdw_e = self.vdw_r = None
        if method == DFIRE:
            # src/dfire.rs:177-183 ; r3_to_numerical panics on unknown residues, ATOMNUMBER on unknown atoms
            ty = []
            for a in atoms:
                key = (a.resname, a.name)
                if key not in t["dfire"]:
                    raise KeyError(f"Not supported atom type {a.resname}{a.name}")
                ty.append(t["dfire"][key])
            self.dfire_type = np.array(ty, dtype=np.int32)
        else:
            amber = t["amber_pydock"] if method == PYDOCK else t["amber"]
            ele = t["ele_pydock"] if method == PYDOCK else t["ele"]
            q, e, r = [], [], []
            for a in atoms:
                # src/dna.rs:314-356 ; src/pydock.rs:318-372
                atom_id = f"{a.resname}-{a.name}"
                if atom_id in amber:
                    at = amber[atom_id]
                elif a.name in ("H1", "H2", "H3"):
                    atom_id = f"{a.resname}-H"
                    at = amber[atom_id]
                elif method == PYDOCK:
                    atom_id = f"*-{a.name[0]}"
                    at = amber[atom_id]
                else:
                    raise KeyError(f"DNA Error: Atom [{atom_id}] not supported")
                q.append(ele[atom_id] if atom_id in ele else t["nt_ele"][atom_id])
                e.append(t["vdw_e"][at])
                r.append(t["vdw_r"][at])
            self.ele = np.array(q, dtype=np.float64)
            self.vdw_e = np.array(e, dtype=np.float64)
            self.vdw_r = np.array(r, dtype=np.float64)
        self.n_modes = int(n_modes)
        if modes is not None and n_modes > 0:
            self.modes = np.ascontiguousarray(modes, dtype=np.float64).reshape(-1)
            assert self.modes.size == n * 3 * n_modes, "ANM size mismatch (src/bin/lightdock-rust.rs:233-235)"
        else:
            self.modes = np.zeros(0, dtype=np.float64)
            self.n_modes = 0 if modes is None else int(n_modes)


# ------------------------------------------------------------------------------------------------
class _CMol(C.Structure):
    _fields_ = [("n_atoms", C.c_int32), ("coords", C.c_void_p), ("dfire_type", C.c_void_p),
                ("ele_charge", C.c_void_p), ("vdw_energy", C.c_void_p), ("vdw_radius", C.c_void_p),
                ("n_modes", C.c_int32), ("modes", C.c_void_p), ("n_restraints", C.c_int32),
                ("rst_offsets", C.c_void_p), ("rst_atoms", C.c_void_p), ("n_membrane", C.c_int32),
                ("membrane", C.c_void_p)]


class _CComplex(C.Structure):
    _fields_ = [("method", C.c_int32), ("use_anm", C.c_int32), ("rec", _CMol), ("lig", _CMol),
                ("dfire_potential", C.c_void_p)]


class Diag(C.Structure):
    _fields_ = [("raw_sum", C.c_double), ("raw_sum2", C.c_double), ("n_in_cutoff", C.c_int64),
                ("n_in_cutoff2", C.c_int64), ("n_interface_pairs", C.c_int64), ("bin_hist", C.c_int64 * 21),
                ("rec_rst_hit", C.c_int32), ("lig_rst_hit", C.c_int32), ("membrane_hit", C.c_int32)]


class GsoStats(C.Structure):
    _fields_ = [("n_energy_calls", C.c_int64)]


def _ptr(a):
    return None if a is None or a.size == 0 else a.ctypes.data


class Complex:
    """The scoring object (`DFIRE` / `DNA` / `PYDOCK` struct of the reference)."""

    def __init__(self, rec: Molecule, lig: Molecule, method, use_anm, potential=None):
        self.lib = _lib()
        self.rec, self.lig, self.method, self.use_anm = rec, lig, method, bool(use_anm)
        self.potential = None
        if method == DFIRE:
            assert potential is not None and potential.size >= 169 * 169 * 20
            self.potential = np.ascontiguousarray(potential[:169 * 169 * 20], dtype=np.float64)
        assert self.lib.oracle_sizeof_complex() == C.sizeof(_CComplex)
        assert self.lib.oracle_sizeof_diag() == C.sizeof(Diag)
        self.c = _CComplex()
        self.c.method = 0 if method == DFIRE else 1
        self.c.use_anm = int(self.use_anm)
        for cm, m in ((self.c.rec, rec), (self.c.lig, lig)):
            cm.n_atoms = m.n
            cm.coords = _ptr(m.coords)
            cm.dfire_type = _ptr(m.dfire_type)
            cm.ele_charge = _ptr(m.ele)
            cm.vdw_energy = _ptr(m.vdw_e)
            cm.vdw_radius = _ptr(m.vdw_r)
            cm.n_modes = m.n_modes if self.use_anm else 0
            cm.modes = _ptr(m.modes)
            cm.n_restraints = len(m.rst_offsets) - 1
            cm.rst_offsets = _ptr(m.rst_offsets)
            cm.rst_atoms = _ptr(m.rst_atoms)
            cm.n_membrane = m.membrane.size
            cm.membrane = _ptr(m.membrane)
        self.c.dfire_potential = _ptr(self.potential)

    @property
    def pose_len(self):
        return 7 + ((self.rec.n_modes + self.lig.n_modes) if self.use_anm else 0)

    def _poses(self, poses):
        poses = np.ascontiguousarray(poses, dtype=np.float64)
        if poses.ndim == 1:
            poses = poses.reshape(1, -1)
        assert poses.shape[1] >= self.pose_len
        return np.ascontiguousarray(poses[:, :self.pose_len])

    def energy(self, poses, detail=False):
        poses = self._poses(poses)
        n = poses.shape[0]
        e = np.zeros(n, dtype=np.float64)
        if not detail:
            rc = self.lib.oracle_score_batch(C.byref(self.c), n, C.c_void_p(poses.ctypes.data),
                                             C.c_void_p(e.ctypes.data), None, None, None, None, None)
            assert rc == 0
            return e
        diag = (Diag * n)()
        irec = np.zeros((n, self.rec.n), dtype=np.uint8)
        ilig = np.zeros((n, self.lig.n), dtype=np.uint8)
        crec = np.zeros((n, self.rec.n, 3), dtype=np.float64)
        clig = np.zeros((n, self.lig.n, 3), dtype=np.float64)
        rc = self.lib.oracle_score_batch(C.byref(self.c), n, C.c_void_p(poses.ctypes.data), C.c_void_p(e.ctypes.data),
                                         diag, C.c_void_p(irec.ctypes.data), C.c_void_p(ilig.ctypes.data),
                                         C.c_void_p(crec.ctypes.data), C.c_void_p(clig.ctypes.data))
        assert rc == 0
        return e, dict(
            raw_sum=np.array([d.raw_sum for d in diag]), raw_sum2=np.array([d.raw_sum2 for d in diag]),
            n_in_cutoff=np.array([d.n_in_cutoff for d in diag], dtype=np.int64),
            n_in_cutoff2=np.array([d.n_in_cutoff2 for d in diag], dtype=np.int64),
            n_interface_pairs=np.array([d.n_interface_pairs for d in diag], dtype=np.int64),
            bin_hist=np.array([list(d.bin_hist) for d in diag], dtype=np.int64),
            rec_rst_hit=np.array([d.rec_rst_hit for d in diag], dtype=np.int32),
            lig_rst_hit=np.array([d.lig_rst_hit for d in diag], dtype=np.int32),
            membrane_hit=np.array([d.membrane_hit for d in diag], dtype=np.int32),
            iface_rec=irec, iface_lig=ilig, coords_rec=crec, coords_lig=clig)

    def energy_mt(self, poses, n_threads):
        poses = self._poses(poses)
        e = np.zeros(poses.shape[0], dtype=np.float64)
        rc = self.lib.oracle_score_batch_mt(C.byref(self.c), poses.shape[0], C.c_void_p(poses.ctypes.data),
                                            C.c_void_p(e.ctypes.data), int(n_threads))
        assert rc == 0
        return e

    def gso_run(self, positions, seed, steps, out_dir=None, trace=False, threads=1):
        """The reference's GSO loop.  threads > 1 scores each step's batch of moved glowworms on that many threads
        (same per-pose function: bit-identical results), so that a 100-step 1k4c trajectory takes seconds."""
        pos = self._poses(positions)
        n = pos.shape[0]
        final = np.zeros_like(pos)
        tr = np.zeros((steps, n, 5 + self.pose_len), dtype=np.float64) if trace else None
        st = GsoStats()
        rc = self.lib.oracle_gso_run_mt(C.byref(self.c), n, C.c_void_p(pos.ctypes.data), C.c_uint64(seed), int(steps),
                                        out_dir.encode() if out_dir else None, C.c_void_p(final.ctypes.data),
                                        C.c_void_p(tr.ctypes.data) if trace else None, C.byref(st), int(threads))
        assert rc == 0, rc
        return final, tr, int(st.n_energy_calls)


# ------------------------------------------------------------------------------------------------
class Rng:
    """rand 0.7.3 `StdRng::seed_from_u64` + `gen::<f64>()` (src/lib.rs:38, src/swarm.rs:118)."""

    def __init__(self, seed):
        self.lib = _lib()
        self.buf = C.create_string_buffer(self.lib.oracle_rng_sizeof())
        self.lib.oracle_rng_seed(self.buf, C.c_uint64(seed))

    def f64(self):
        return self.lib.oracle_rng_f64(self.buf)


def rotate(q, v):
    out = (C.c_double * 3)()
    _lib().oracle_rotate((C.c_double * 4)(*q), (C.c_double * 3)(*v), out)
    return list(out)


def slerp(a, b, t):
    out = (C.c_double * 4)()
    _lib().oracle_slerp((C.c_double * 4)(*a), (C.c_double * 4)(*b), C.c_double(t), out)
    return list(out)


# ------------------------------------------------------------------------------------------------
# DFIRE potential table
def load_dcparams(path):
    """src/dfire.rs:236-257: first 169*169*20 lines, one f64 per line."""
    vals = []
    with open(path) as f:
        for line in f:
            vals.append(float(line.strip()))
            if len(vals) == 169 * 169 * 20:
                break
    return np.array(vals, dtype=np.float64)


def synthetic_dcparams(seed=20240324):
    """Seeded stand-in for the missing data/DCparams (same shape/scale as the real table as far as
    the commented test src/dfire.rs:370-380 reveals: 10.0 at the shortest bins, O(1) values after,
    0.0 tail).  Values are rounded to 9 decimals so the text round-trip is exact."""
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
        for v in table:
            f.write(f"{v:.9f}\n")


def real_or_synthetic_dcparams():
    """Returns (table, kind).  A real DCparams is used when LIGHTDOCK_DATA points at one."""
    d = os.environ.get("LIGHTDOCK_DATA")
    if d and os.path.exists(os.path.join(d, "DCparams")):
        return load_dcparams(os.path.join(d, "DCparams")), "real"
    return synthetic_dcparams(), "synthetic"


# ------------------------------------------------------------------------------------------------
# whole-case loader mirroring `simulate` (src/bin/lightdock-rust.rs:158-333)
def load_case(case_dir, method, setup_name="setup.json", positions=None, potential=None):
    import json
    with open(os.path.join(case_dir, setup_name)) as f:
        setup = json.load(f)
    use_anm = bool(setup["use_anm"])
    rec_atoms = read_pdb(os.path.join(case_dir, "lightdock_" + setup["receptor_pdb"]))
    lig_atoms = read_pdb(os.path.join(case_dir, "lightdock_" + setup["ligand_pdb"]))
    rec_nm = lig_nm = None
    if use_anm:
        if setup["anm_rec"] > 0:
            rec_nm = np.load(os.path.join(case_dir, "rec_nm.npy"))
        if setup["anm_lig"] > 0:
            lig_nm = np.load(os.path.join(case_dir, "lig_nm.npy"))
    ra = (setup.get("receptor_restraints") or {}).get("active", [])
    la = (setup.get("ligand_restraints") or {}).get("active", [])
    rec = Molecule(rec_atoms, method, ra, rec_nm, setup["anm_rec"])
    lig = Molecule(lig_atoms, method, la, lig_nm, setup["anm_lig"])
    if method == DFIRE and potential is None:
        potential, _ = real_or_synthetic_dcparams()
    cx = Complex(rec, lig, method, use_anm, potential)
    if positions is None:
        p = os.path.join(case_dir, "initial_positions_0.dat")
        if not os.path.exists(p):
            p = os.path.join(case_dir, "init", "initial_positions_0.dat")
        positions = p
    pos = np.array([[float(x) for x in line.split(" ")] for line in open(positions).read().splitlines() if line])
    seed = setup.get("seed") or 324324
    return cx, pos, seed, setup


def parse_gso_out(path):
    """-> (poses [n, k], luciferin, n_neighbors, vision, scoring)"""
    poses, luc, nn, vis, sc = [], [], [], [], []
    for line in open(path):
        if line.startswith("#"):
            continue
        a, b = line.split(")")
        poses.append([float(x) for x in a.strip("( ").split(",")])
        f = b.split()
        luc.append(float(f[2])); nn.append(int(f[3])); vis.append(float(f[4])); sc.append(float(f[5]))
    return np.array(poses), np.array(luc), np.array(nn), np.array(vis), np.array(sc)
