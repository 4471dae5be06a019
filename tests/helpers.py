"""Shared test helpers: build the product's Scorer from the oracle's numeric model and compare."""
import functools
import os

import numpy as np

import oracle as O  # test infrastructure (oracle/oracle.py)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# tolerance stated by BASELINE.json north_star: per-pose energies within 1e-6 relative of the f64 reference
ENERGY_RTOL = 1e-6


def mol_kwargs(m, use_anm):
    return dict(coords=m.coords, dfire_type=m.dfire_type, ele_charge=m.ele, vdw_energy=m.vdw_e, vdw_radius=m.vdw_r,
                modes=m.modes if (use_anm and m.n_modes > 0) else None, n_modes=m.n_modes if use_anm else 0,
                rst_offsets=m.rst_offsets, rst_atoms=m.rst_atoms, membrane=m.membrane)


def scorer_from_oracle(cx, device=0):
    import ldb200
    method = {O.DFIRE: ldb200.METHOD_DFIRE, O.DNA: ldb200.METHOD_DNA, O.PYDOCK: ldb200.METHOD_PYDOCK}[cx.method]
    return ldb200.Scorer(method, mol_kwargs(cx.rec, cx.use_anm), mol_kwargs(cx.lig, cx.use_anm), use_anm=cx.use_anm,
                         dfire_potential=cx.potential, device=device)


@functools.lru_cache(maxsize=None)
def case(name, method):
    """(oracle Complex, start positions, seed) of a golden example directory."""
    cx, pos, seed, _ = O.load_case(os.path.join(GOLDEN, name), method)
    return cx, pos, seed


def random_poses(rng, n, pose_len, centre=(0, 0, 0), spread=8.0, ext_scale=3.0):
    t = np.asarray(centre) + rng.normal(0, spread, size=(n, 3))
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    ext = rng.normal(0, ext_scale, size=(n, pose_len - 7))
    return np.hstack([t, q, ext])


def assert_parity(e_gpu, d_gpu, e_ref, d_ref, method, rtol=ENERGY_RTOL):
    """Integer outputs bit-exact; energies within rtol relative."""
    for k in ("n_in_cutoff", "n_interface_pairs", "rec_rst_hit", "lig_rst_hit", "membrane_hit"):
        np.testing.assert_array_equal(d_gpu[k], d_ref[k], err_msg=k)
    if method == O.DFIRE:
        np.testing.assert_array_equal(d_gpu["bin_hist"], d_ref["bin_hist"], err_msg="bin histogram")
    else:
        np.testing.assert_array_equal(d_gpu["n_in_cutoff2"], d_ref["n_in_cutoff2"], err_msg="n_in_cutoff2")
    np.testing.assert_array_equal(d_gpu["iface_rec"], d_ref["iface_rec"], err_msg="receptor interface flags")
    np.testing.assert_array_equal(d_gpu["iface_lig"], d_ref["iface_lig"], err_msg="ligand interface flags")
    scale = np.maximum(np.abs(e_ref), 1e-300)
    rel = np.abs(e_gpu - e_ref) / scale
    assert rel.max() <= rtol, f"max relative energy error {rel.max():.3e} > {rtol}"
    return rel.max()
