"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): distance-bin indices, atom-type lookups and interface-contact counts
bit-exact; per-pose energies within 1e-6 relative (ENERGY_RTOL in helpers.py).
DFIRE runs on the seeded synthetic DCparams unless LIGHTDOCK_DATA points at a real one (the real
table is absent from the build container: DFIRE parity against the reference's own numbers is unpinned).
"""
import numpy as np
import pytest

import oracle as O
import ldb200
from helpers import ENERGY_RTOL, assert_parity, case, random_poses
from helpers import scorer_from_oracle as _scorer_auto

pytestmark = pytest.mark.gpu


def scorer_from_oracle(cx):
    """This module pins the GENERIC pair kernels (per-pose ligand transform + sphere culling; the only path
    for DNA/pyDock and for ligands with ANM modes).  tests/test_gpu_rigid_path.py covers the rigid-ligand
    DFIRE kernel that AUTO selects where it applies, with the same cases."""
    sc = _scorer_auto(cx)
    sc.set_path(ldb200.PATH_GENERIC)
    return sc


CASES = [("1azp", O.DNA), ("1azp", O.PYDOCK), ("1czy", O.DFIRE), ("1ppe", O.DFIRE), ("2uuy", O.DFIRE),
         ("1k4c", O.DFIRE), ("ab_icode", O.DFIRE)]


@pytest.mark.parametrize("name,method", CASES)
def test_transform_bit_exact(name, method):
    cx, pos, _ = case(name, method)
    sc = scorer_from_oracle(cx)
    poses = pos[:16]
    rec, lig = sc.transform(poses)
    _, d = cx.energy(poses, detail=True)
    assert np.array_equal(rec, d["coords_rec"]), "receptor coordinates differ from the oracle"
    assert np.array_equal(lig, d["coords_lig"]), "ligand coordinates differ from the oracle"


@pytest.mark.parametrize("name,method", CASES)
def test_start_positions_parity(name, method):
    cx, pos, _ = case(name, method)
    sc = scorer_from_oracle(cx)
    n = 200 if cx.rec.n * cx.lig.n < 2_000_000 else 24  # keep the oracle side to seconds
    poses = pos[:n]
    e_gpu, d_gpu = sc.energy_detail(poses)
    e_ref, d_ref = cx.energy(poses, detail=True)
    assert_parity(e_gpu, d_gpu, e_ref, d_ref, cx.method)
    # the plain (non-detail) kernel instantiation must give the same bits as the detail one
    assert np.array_equal(sc.energy(poses), e_gpu)


@pytest.mark.parametrize("name,method", [("1azp", O.DNA), ("2uuy", O.DFIRE), ("1k4c", O.DFIRE)])
def test_random_close_poses_parity(name, method):
    """Poses pushed into the receptor: many contacts, restraints and membrane beads hit."""
    cx, pos, _ = case(name, method)
    sc = scorer_from_oracle(cx)
    rng = np.random.default_rng(7)
    centre = cx.rec.coords.mean(axis=0)
    n = 48 if cx.rec.n * cx.lig.n < 2_000_000 else 12
    poses = random_poses(rng, n, cx.pose_len, centre=centre, spread=12.0)
    e_gpu, d_gpu = sc.energy_detail(poses)
    e_ref, d_ref = cx.energy(poses, detail=True)
    assert_parity(e_gpu, d_gpu, e_ref, d_ref, cx.method)
    assert d_ref["n_interface_pairs"].max() > 0


def test_dna_known_answer(golden_dir):
    """src/dna.rs:538-572 / src/pydock.rs:553-587: identity pose on tests/1azp -> -364.88126358158974."""
    g = golden_dir + "/unit/1azp/"
    for method in (O.DNA, O.PYDOCK):
        rec = O.Molecule(O.read_pdb(g + "1azp_receptor.pdb"), method)
        lig = O.Molecule(O.read_pdb(g + "1azp_ligand.pdb"), method)
        cx = O.Complex(rec, lig, method, False)
        sc = scorer_from_oracle(cx)
        e = sc.energy([[0, 0, 0, 1, 0, 0, 0]])[0]
        assert abs(e - (-364.88126358158974)) <= ENERGY_RTOL * 364.88126358158974


def test_1azp_gso1_golden(golden_dir):
    """All 200 (pose -> energy) pairs of example/1azp/swarm_0/gso_1.out (DNA + ANM + restraints)."""
    cx, pos, _ = case("1azp", O.DNA)
    sc = scorer_from_oracle(cx)
    _, _, _, _, score = O.parse_gso_out(golden_dir + "/1azp/swarm_0/gso_1.out")
    e = sc.energy(pos)
    # half a unit of the 8-decimal print precision + 1e-11 relative (the kernel's reciprocals carry 1e-12; the
    # north-star tolerance is 1e-6)
    assert (np.abs(e - score) <= 5.0e-9 + 1e-11 * np.abs(score)).all()
    # printed to 8 decimals: a value within 1e-10 of a rounding boundary may print one unit away
    assert (np.abs(np.round(e, 8) - score) <= 1.0e-8 * (1 + 1e-6)).all()
    assert (np.round(e, 8) == score).mean() >= 0.9


def test_rec_splits_same_result():
    """Splitting the receptor over several CTAs per pose must not change a single bit: per-tile sums
    are combined in tile order whatever the split (batch-invariant energies)."""
    cx, pos, _ = case("1k4c", O.DFIRE)
    sc = scorer_from_oracle(cx)
    poses = pos[:32]
    sc.set_rec_splits(1)
    e1, d1 = sc.energy_detail(poses)
    for s in (2, 5, 27):
        sc.set_rec_splits(s)
        e2, d2 = sc.energy_detail(poses)
        np.testing.assert_array_equal(d1["bin_hist"], d2["bin_hist"])
        np.testing.assert_array_equal(d1["iface_lig"], d2["iface_lig"])
        np.testing.assert_array_equal(d1["iface_rec"], d2["iface_rec"])
        assert np.array_equal(e1, e2)


def test_run_to_run_deterministic():
    cx, pos, _ = case("1k4c", O.DFIRE)
    sc = scorer_from_oracle(cx)
    e1 = sc.energy(pos)
    e2 = sc.energy(pos)
    assert np.array_equal(e1, e2)


def test_edge_cases():
    cx, pos, _ = case("1ppe", O.DFIRE)
    sc = scorer_from_oracle(cx)
    assert sc.energy(np.zeros((0, 7))).shape == (0,)
    # a pose far away: no pair in the cut-off -> (0*0.0157 - 4.7) * -1 = 4.7 exactly (src/dfire.rs:347)
    far = np.array([[1e4, 0, 0, 1, 0, 0, 0]], dtype=np.float64)
    e, d = sc.energy_detail(far)
    assert d["n_in_cutoff"][0] == 0 and e[0] == 4.7
    # un-normalised quaternion: rotate() divides by norm2 (src/qt.rs:48-50), must match the oracle
    p = pos[:8].copy()
    p[:, 3:7] *= 1.7
    e_gpu, d_gpu = sc.energy_detail(p)
    e_ref, d_ref = cx.energy(p, detail=True)
    assert_parity(e_gpu, d_gpu, e_ref, d_ref, cx.method)


def _threshold_complex():
    """One receptor atom at the origin, ligand atoms on the x axis at the decision thresholds of the DFIRE
    loop and a few ulps / small offsets either side: every bin edge (k+1)/2, the 15 A cut-off (dist == 225
    gives DIST_TO_BINS[29]-1 = 20, i.e. the next ligand type's bin 0) and the 2.45 A interface edge."""
    edges = [(k + 1) / 2.0 for k in range(0, 30)] + [2.45, 15.0]
    xs = []
    for e in edges:
        v = np.float64(e)
        xs += [v, np.nextafter(v, 0), np.nextafter(v, 100), np.nextafter(np.nextafter(v, 0), 0),
               np.nextafter(np.nextafter(v, 100), 100), v - 1e-7, v + 1e-7, v - 3e-5, v + 3e-5, v - 1e-3, v + 1e-3]
    xs = np.array(xs, dtype=np.float64)
    lig = np.zeros((xs.size, 3)); lig[:, 0] = xs
    rec = np.zeros((3, 3)); rec[1] = [0, 400, 0]; rec[2] = [0, 0, -400]

    class M:  # duck-typed oracle Molecule
        pass
    def mol(c, types):
        m = M()
        m.n = len(c); m.coords = np.ascontiguousarray(c); m.dfire_type = np.asarray(types, np.int32)
        m.ele = m.vdw_e = m.vdw_r = None
        m.membrane = np.zeros(0, np.int32); m.rst_offsets = np.array([0, 1], np.int32); m.rst_atoms = np.array([0], np.int32)
        m.n_modes = 0; m.modes = np.zeros(0)
        return m
    rng = np.random.default_rng(11)
    return mol(rec, [5, 17, 167]), mol(lig, rng.integers(0, 168, size=xs.size))


def test_dfire_decision_thresholds_exact():
    """Pairs sitting on / next to every decision threshold must fall on the reference's side of it."""
    rec, lig = _threshold_complex()
    pot, _ = O.real_or_synthetic_dcparams()
    cx = O.Complex(rec, lig, O.DFIRE, False, pot)
    sc = scorer_from_oracle(cx)
    poses = np.array([[0, 0, 0, 1, 0, 0, 0],            # identity: coordinates exact
                      [0, 0, 0, 0, 1, 0, 0],            # 180 deg about x: x unchanged
                      [1e-9, 0, 0, 1, 0, 0, 0], [-1e-9, 0, 0, 1, 0, 0, 0],
                      [3e-6, 0, 0, 1, 0, 0, 0], [0, 2e-4, 0, 1, 0, 0, 0],
                      [0.25, 0, 0, 1, 0, 0, 0], [-0.25, 0, 0, 1, 0, 0, 0]], dtype=np.float64)
    e_gpu, d_gpu = sc.energy_detail(poses)
    e_ref, d_ref = cx.energy(poses, detail=True)
    assert_parity(e_gpu, d_gpu, e_ref, d_ref, O.DFIRE)
    assert d_ref["bin_hist"][0][20] >= 1, "dist == 225 exactly must land in bin 20 (src/dfire.rs:337)"
    assert d_gpu["n_exact_fallback"][0] > 0, "threshold pairs must take the exact FP64 path"


def test_dfire_exact_fallback_is_rare():
    cx, pos, _ = case("1k4c", O.DFIRE)
    sc = scorer_from_oracle(cx)
    _, d = sc.energy_detail(pos[:64])
    frac = d["n_exact_fallback"].sum() / d["n_in_cutoff"].sum()
    assert frac < 0.01, frac


def test_large_coordinates_stay_exact():
    """f32 margins scale with the coordinate magnitude: shift the whole system 5000 A away from the origin."""
    cx, pos, _ = case("1ppe", O.DFIRE)
    shift = np.array([5000.0, -3000.0, 2000.0])
    rec = cx.rec; lig = cx.lig
    import copy
    rec2 = copy.copy(rec); rec2.coords = rec.coords + shift
    cx2 = O.Complex(rec2, lig, O.DFIRE, False, cx.potential)
    sc = scorer_from_oracle(cx2)
    p = pos[:24].copy(); p[:, :3] += shift
    e_gpu, d_gpu = sc.energy_detail(p)
    e_ref, d_ref = cx2.energy(p, detail=True)
    assert_parity(e_gpu, d_gpu, e_ref, d_ref, O.DFIRE)


def _dna_threshold_points():
    """Ligand positions (receptor atom at the origin) whose squared distance sits on / within a few ulp of the
    three DNA cut-offs (900, 100, 3.9*3.9; src/dna.rs:17-18, src/constants.rs:15), including points where the
    reference's never-fused d2 = (x*x + y*y) + z*z and a fused evaluation fall on DIFFERENT sides of the
    threshold: the kernel evaluates d2 with FMAs and must hand exactly those pairs to its exact path."""
    from fractions import Fraction as F
    rng = np.random.default_rng(2024)
    pts, straddle = [], 0
    for thr in (900.0, 100.0, 3.9 * 3.9):
        found = 0
        for _ in range(400):
            u = rng.normal(size=3)
            u /= np.linalg.norm(u)
            v = u * np.sqrt(thr)
            # walk x by ulps until the unfused d2 crosses the threshold, then look at the neighbours
            def unfused(p):
                return (p[0] * p[0] + p[1] * p[1]) + p[2] * p[2]
            def fused(p):
                a = float(F(p[0]) * F(p[0]))
                a = float(F(p[1]) * F(p[1]) + F(a))
                return float(F(p[2]) * F(p[2]) + F(a))
            p = v.copy()
            step = np.sign(p[0]) if p[0] != 0 else 1.0
            for _ in range(64):  # move |x| towards the crossing
                if unfused(p) <= thr:
                    q = p.copy(); q[0] = np.nextafter(q[0], q[0] + step)
                    if unfused(q) > thr:
                        break
                    p = q
                else:
                    p[0] = np.nextafter(p[0], 0.0)
            for k in range(-2, 3):
                q = p.copy()
                for _ in range(abs(k)):
                    q[0] = np.nextafter(q[0], q[0] + (step if k > 0 else -step))
                if (unfused(q) <= thr) != (fused(q) <= thr):
                    pts.append(q); straddle += 1; found += 1
                elif found < 12 and abs(k) <= 1:
                    pts.append(q)
            if found >= 12:
                break
        # the threshold itself, exactly, on an axis
        r = np.sqrt(thr)
        if r * r == thr:
            pts.append(np.array([r, 0.0, 0.0]))
    return np.array(pts), straddle


def test_dna_decision_thresholds_exact():
    """Pairs on / next to the three DNA cut-offs must be counted exactly as the reference's never-fused d2 decides,
    also where a fused d2 would decide otherwise (the kernel's d2 is fused; its high-word test must catch them)."""
    pts, straddle = _dna_threshold_points()
    assert straddle >= 6, "the generator must produce pairs whose fused and unfused d2 disagree about a cut-off"

    class M:
        pass
    def mol(c, q):
        m = M()
        m.n = len(c); m.coords = np.ascontiguousarray(c, dtype=np.float64); m.dfire_type = None
        m.ele = np.asarray(q, np.float64); m.vdw_e = np.full(m.n, 0.1094); m.vdw_r = np.full(m.n, 1.908)
        m.membrane = np.zeros(0, np.int32); m.rst_offsets = np.array([0, 1], np.int32); m.rst_atoms = np.array([0], np.int32)
        m.n_modes = 0; m.modes = np.zeros(0)
        return m
    rng = np.random.default_rng(5)
    rec = mol(np.zeros((1, 3)), [0.7])
    lig = mol(pts, rng.uniform(-0.8, 0.8, size=len(pts)))
    cx = O.Complex(rec, lig, O.DNA, False)
    sc = scorer_from_oracle(cx)
    poses = np.array([[0, 0, 0, 1, 0, 0, 0], [0, 0, 0, -1, 0, 0, 0]], dtype=np.float64)
    e_gpu, d_gpu = sc.energy_detail(poses)
    e_ref, d_ref = cx.energy(poses, detail=True)
    assert_parity(e_gpu, d_gpu, e_ref, d_ref, O.DNA)
    assert np.array_equal(sc.energy(poses), e_gpu)
    assert 0 < d_ref["n_in_cutoff"][0] < len(pts) and 0 < d_ref["n_in_cutoff2"][0] < d_ref["n_in_cutoff"][0]
    assert 0 < d_ref["n_interface_pairs"][0] < d_ref["n_in_cutoff2"][0]


def test_dna_clamp_and_close_contacts():
    """Coulomb clamp (+-4/332, src/dna.rs:484-489) and the capped 12-6 term on very close pairs: the kernel takes
    its clamped tile form there; energies must stay within tolerance and counts exact."""
    cx, pos, _ = case("1azp", O.DNA)
    sc = scorer_from_oracle(cx)
    rng = np.random.default_rng(17)
    centre = cx.rec.coords.mean(axis=0)
    poses = random_poses(rng, 32, cx.pose_len, centre=centre, spread=4.0, ext_scale=1.0)  # ligand inside the receptor
    e_gpu, d_gpu = sc.energy_detail(poses)
    e_ref, d_ref = cx.energy(poses, detail=True)
    rel = assert_parity(e_gpu, d_gpu, e_ref, d_ref, O.DNA)
    assert d_ref["n_interface_pairs"].min() > 100
    assert rel < 1e-9, rel


def test_dna_many_vdw_types_takes_the_per_atom_form():
    """More than 1024 distinct (vdw energy, radius) pairs: the kernel cannot tabulate the 12-6 term by type pair and
    evaluates it from the per-atom parameters (the TAB == false instance of dna_tile); same parity bar."""
    import copy
    cx, pos, _ = case("1azp", O.DNA)
    rng = np.random.default_rng(23)
    rec, lig = copy.copy(cx.rec), copy.copy(cx.lig)
    rec.vdw_e = cx.rec.vdw_e * rng.uniform(0.9, 1.1, size=cx.rec.n)
    lig.vdw_r = cx.lig.vdw_r * rng.uniform(0.95, 1.05, size=cx.lig.n)
    cx2 = O.Complex(rec, lig, O.DNA, cx.use_anm)
    sc = scorer_from_oracle(cx2)
    centre = cx.rec.coords.mean(axis=0)
    poses = np.vstack([pos[:24], random_poses(rng, 24, cx.pose_len, centre=centre, spread=10.0, ext_scale=1.0)])
    e_gpu, d_gpu = sc.energy_detail(poses)
    e_ref, d_ref = cx2.energy(poses, detail=True)
    assert_parity(e_gpu, d_gpu, e_ref, d_ref, O.DNA)
    assert np.array_equal(sc.energy(poses), e_gpu)


@pytest.mark.parametrize("method", [O.DNA, O.PYDOCK])
def test_dna_fused_transform_equals_the_transform_kernel(method):
    """DNA/pyDock: the pair kernel transforms its pose itself (ligand in the prologue, receptor ANM per tile); with
    ld_set_option("dna_fused", 0) the poses go through transform_kernel and per-pose coordinate blocks instead.  Same
    operations in the same order: every output identical, energies BIT FOR BIT -- large batch, one swarm (receptor
    splits) and single poses."""
    import ldb200
    cx, pos, _ = case("1azp", method)
    rng = np.random.default_rng(23)
    poses = np.tile(pos, (6, 1))
    poses[:, :3] += rng.normal(0, 1.5, size=(len(poses), 3))
    poses[:, 7:] *= rng.uniform(0.5, 1.5, size=(len(poses), 1))
    fused = scorer_from_oracle(cx)
    ldb200.set_option("dna_fused", 0)
    try:
        plain = scorer_from_oracle(cx)  # the option is read by ld_create
    finally:
        ldb200.set_option("dna_fused", 1)
    e_p, d_p = plain.energy_detail(poses)
    e_p200 = plain.energy(poses[:200])
    e_f, d_f = fused.energy_detail(poses)
    for k in d_f:
        np.testing.assert_array_equal(d_f[k], d_p[k], err_msg=k)
    assert np.array_equal(e_f, e_p)
    assert np.array_equal(fused.energy(poses), e_f)
    assert np.array_equal(plain.energy(poses), e_p)
    assert plain.stats()["kernel_launches"] == 3 and fused.stats()["kernel_launches"] == 2
    assert np.array_equal(fused.energy(poses[:200]), e_p200)
    assert np.array_equal(np.array([fused.energy(poses[i:i + 1])[0] for i in range(8)]), e_f[:8])
