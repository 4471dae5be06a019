"""ctypes binding of the C ABI in include/lightdock_b200.h (liblightdock_b200.so).

This is the thin Python face of the product used by the tests, bench.py and smoke(): it only
marshals numpy arrays into `ld_complex_desc` and forwards to the CUDA library.  It never computes
energies itself and never touches oracle/; if the shared library is missing it raises.
"""
import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG_DIR, "liblightdock_b200.so")

METHOD_DFIRE, METHOD_DNA, METHOD_PYDOCK = 0, 1, 2
DFIRE_TABLE_LEN = 169 * 169 * 20

EXPORTS = ["ld_create", "ld_destroy", "ld_pose_len", "ld_score_batch", "ld_score_batch_device",
           "ld_score_batch_detail", "ld_transform_batch", "ld_get_stats", "ld_set_rec_splits",
           "ld_set_profiling", "ld_probe_peaks", "ld_last_error", "ld_version", "ld_set_path", "ld_path_info", "ld_device_count", "ld_score_batch_begin", "ld_score_batch_end", "ld_get_stats_slot",
           "ld_set_option", "ld_init_device", "ld_get_create_ms",
           "ld_gso_create", "ld_gso_run", "ld_gso_state", "ld_gso_steps", "ld_gso_energy_calls", "ld_gso_destroy"]

PATH_AUTO, PATH_GENERIC, PATH_RIGID = 0, 1, 2


class MoleculeDesc(C.Structure):
    _fields_ = [("n_atoms", C.c_int32), ("coords", C.c_void_p), ("dfire_type", C.c_void_p),
                ("ele_charge", C.c_void_p), ("vdw_energy", C.c_void_p), ("vdw_radius", C.c_void_p),
                ("n_modes", C.c_int32), ("modes", C.c_void_p), ("n_restraints", C.c_int32),
                ("rst_offsets", C.c_void_p), ("rst_atoms", C.c_void_p), ("n_membrane", C.c_int32),
                ("membrane", C.c_void_p)]


class ComplexDesc(C.Structure):
    _fields_ = [("method", C.c_int32), ("use_anm", C.c_int32), ("receptor", MoleculeDesc),
                ("ligand", MoleculeDesc), ("dfire_potential", C.c_void_p), ("device", C.c_int32),
                ("reserved", C.c_int32)]


class PoseDetail(C.Structure):
    _fields_ = [("raw_sum", C.c_double), ("raw_sum2", C.c_double), ("n_in_cutoff", C.c_int64),
                ("n_in_cutoff2", C.c_int64), ("n_interface_pairs", C.c_int64), ("bin_hist", C.c_int64 * 21),
                ("rec_rst_hit", C.c_int32), ("lig_rst_hit", C.c_int32), ("membrane_hit", C.c_int32),
                ("reserved", C.c_int32), ("n_pairs_tested", C.c_int64), ("n_exact_fallback", C.c_int64)]


class BatchStats(C.Structure):
    _fields_ = [("n_poses", C.c_int64), ("pair_evals_bruteforce", C.c_int64), ("kernel_launches", C.c_int32),
                ("rec_splits", C.c_int32), ("device_ms", C.c_double), ("transform_ms", C.c_double),
                ("pair_ms", C.c_double), ("finalize_ms", C.c_double), ("path", C.c_int32),
                ("pair_launches", C.c_int32)]


class LdError(RuntimeError):
    pass


_lib = None


def load_library():
    """Loads the CUDA library; raises (no fallback) if it has not been built."""
    global _lib
    if _lib is None:
        path = os.environ.get("LDB200_LIB", LIB_PATH)  # kernel A/B experiments load an alternative build
        if not os.path.exists(path):
            raise LdError(f"{path} not built: run `make -C lightdock-rust_b200` (or __graft_entry__.build())")
        lib = C.CDLL(path)
        lib.ld_last_error.restype = C.c_char_p
        lib.ld_version.restype = C.c_char_p
        lib.ld_create.argtypes = [C.POINTER(ComplexDesc), C.POINTER(C.c_void_p)]
        lib.ld_destroy.argtypes = [C.c_void_p]
        lib.ld_pose_len.argtypes = [C.c_void_p]
        lib.ld_score_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        lib.ld_score_batch_device.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ld_score_batch_detail.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p]
        lib.ld_transform_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ld_get_stats.argtypes = [C.c_void_p, C.POINTER(BatchStats)]
        lib.ld_set_rec_splits.argtypes = [C.c_void_p, C.c_int32]
        lib.ld_set_profiling.argtypes = [C.c_void_p, C.c_int32]
        lib.ld_score_batch_begin.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]
        lib.ld_score_batch_end.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        lib.ld_set_path.argtypes = [C.c_void_p, C.c_int32]
        lib.ld_path_info.argtypes = [C.c_void_p]
        lib.ld_path_info.restype = C.c_char_p
        lib.ld_probe_peaks.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ld_get_stats_slot.argtypes = [C.c_void_p, C.c_int32, C.POINTER(BatchStats)]
        lib.ld_set_option.argtypes = [C.c_char_p, C.c_double]
        lib.ld_init_device.argtypes = [C.c_int32]
        lib.ld_get_create_ms.argtypes = [C.c_void_p, C.c_void_p]
        lib.ld_gso_create.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        lib.ld_gso_run.argtypes = [C.c_void_p, C.c_int32]
        lib.ld_gso_state.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        lib.ld_gso_steps.argtypes = [C.c_void_p]
        lib.ld_gso_steps.restype = C.c_int32
        lib.ld_gso_energy_calls.argtypes = [C.c_void_p]
        lib.ld_gso_energy_calls.restype = C.c_int64
        lib.ld_gso_destroy.argtypes = [C.c_void_p]
        # experiment knobs of the tools (tools/units_sweep.py, ...): forwarded through the API, the library itself
        # never reads the environment
        for env, key in (("LDB200_ROWS", "rigid_rows"), ("LDB200_CELL", "cell_size"),
                         ("LDB200_UNITS_PER_SM", "units_per_sm"), ("LDB200_FLEX_MIN_WARPS", "flex_min_warps"), ("LDB200_COMPACT_TILES", "compact_tiles"), ("LDB200_DNA_FUSED", "dna_fused")):
            if os.environ.get(env):
                _check(lib, lib.ld_set_option(key.encode(), float(os.environ[env])))
        if os.environ.get("LDB200_PATH") == "generic":
            _check(lib, lib.ld_set_option(b"default_path", float(PATH_GENERIC)))
        _lib = lib
    return _lib


def _check(lib, rc):
    if rc != 0:
        raise LdError(f"lightdock_b200 error {rc}: {lib.ld_last_error().decode()}")


def _arr(a, dtype):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def _p(a):
    return None if a is None or a.size == 0 else a.ctypes.data


def molecule_desc(keep, coords, dfire_type=None, ele_charge=None, vdw_energy=None, vdw_radius=None, modes=None,
                  n_modes=0, rst_offsets=None, rst_atoms=None, membrane=None):
    """Builds an ld_molecule_desc; `keep` collects the arrays that must outlive the call."""
    coords = _arr(coords, np.float64).reshape(-1, 3)
    m = MoleculeDesc()
    m.n_atoms = coords.shape[0]
    arrs = dict(coords=coords, dfire_type=_arr(dfire_type, np.int32), ele_charge=_arr(ele_charge, np.float64),
                vdw_energy=_arr(vdw_energy, np.float64), vdw_radius=_arr(vdw_radius, np.float64),
                modes=_arr(modes, np.float64), rst_offsets=_arr(rst_offsets, np.int32),
                rst_atoms=_arr(rst_atoms, np.int32), membrane=_arr(membrane, np.int32))
    for k, v in arrs.items():
        setattr(m, k, _p(v))
        keep.append(v)
    m.n_modes = int(n_modes)
    m.n_restraints = 0 if arrs["rst_offsets"] is None else max(0, arrs["rst_offsets"].size - 1)
    m.n_membrane = 0 if arrs["membrane"] is None else arrs["membrane"].size
    return m


class Scorer:
    """Device-resident scoring object: the `Box<dyn Score>` of src/bin/lightdock-rust.rs:276-316."""

    def __init__(self, method, receptor: dict, ligand: dict, use_anm=False, dfire_potential=None, device=0):
        self.lib = load_library()
        keep = []
        d = ComplexDesc()
        d.method = int(method)
        d.use_anm = int(bool(use_anm))
        d.receptor = molecule_desc(keep, **receptor)
        d.ligand = molecule_desc(keep, **ligand)
        pot = _arr(dfire_potential, np.float64)
        if pot is not None:
            assert pot.size >= DFIRE_TABLE_LEN
        keep.append(pot)
        d.dfire_potential = _p(pot)
        d.device = int(device)
        self.h = C.c_void_p()
        _check(self.lib, self.lib.ld_create(C.byref(d), C.byref(self.h)))
        self.n_rec, self.n_lig = d.receptor.n_atoms, d.ligand.n_atoms
        self.pose_len = self.lib.ld_pose_len(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.ld_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _poses(self, poses):
        poses = np.ascontiguousarray(poses, dtype=np.float64)
        if poses.ndim == 1:
            poses = poses.reshape(1, -1)
        if poses.shape[1] < self.pose_len:
            raise LdError(f"pose rows have {poses.shape[1]} columns, need {self.pose_len}")
        return np.ascontiguousarray(poses[:, :self.pose_len])

    def energy(self, poses):
        """Score::energy for every row of `poses` (host buffers, one batched call)."""
        poses = self._poses(poses)
        e = np.empty(poses.shape[0], dtype=np.float64)
        _check(self.lib, self.lib.ld_score_batch(self.h, poses.shape[0], _p(poses), _p(e)))
        return e

    def energy_begin(self, slot, poses):
        """Enqueue a batch on `slot` (0 or 1) and return at once; energy_end(slot) collects it."""
        poses = self._poses(poses)
        _check(self.lib, self.lib.ld_score_batch_begin(self.h, int(slot), poses.shape[0], _p(poses)))
        return poses.shape[0]

    def energy_end(self, slot, n):
        e = np.empty(n, dtype=np.float64)
        _check(self.lib, self.lib.ld_score_batch_end(self.h, int(slot), _p(e) if n else None))
        return e

    def energy_detail(self, poses):
        poses = self._poses(poses)
        n = poses.shape[0]
        e = np.empty(n, dtype=np.float64)
        det = (PoseDetail * max(n, 1))()
        irec = np.zeros((n, self.n_rec), dtype=np.uint8)
        ilig = np.zeros((n, self.n_lig), dtype=np.uint8)
        _check(self.lib, self.lib.ld_score_batch_detail(self.h, n, _p(poses), _p(e), C.addressof(det), _p(irec),
                                                        _p(ilig)))
        det = det[:n]
        return e, dict(
            raw_sum=np.array([d.raw_sum for d in det]), raw_sum2=np.array([d.raw_sum2 for d in det]),
            n_in_cutoff=np.array([d.n_in_cutoff for d in det], dtype=np.int64),
            n_in_cutoff2=np.array([d.n_in_cutoff2 for d in det], dtype=np.int64),
            n_interface_pairs=np.array([d.n_interface_pairs for d in det], dtype=np.int64),
            bin_hist=np.array([list(d.bin_hist) for d in det], dtype=np.int64).reshape(n, 21),
            rec_rst_hit=np.array([d.rec_rst_hit for d in det], dtype=np.int32),
            lig_rst_hit=np.array([d.lig_rst_hit for d in det], dtype=np.int32),
            membrane_hit=np.array([d.membrane_hit for d in det], dtype=np.int32),
            n_pairs_tested=np.array([d.n_pairs_tested for d in det], dtype=np.int64),
            n_exact_fallback=np.array([d.n_exact_fallback for d in det], dtype=np.int64),
            iface_rec=irec, iface_lig=ilig)

    def transform(self, poses):
        poses = self._poses(poses)
        n = poses.shape[0]
        rec = np.empty((n, self.n_rec, 3), dtype=np.float64)
        lig = np.empty((n, self.n_lig, 3), dtype=np.float64)
        _check(self.lib, self.lib.ld_transform_batch(self.h, n, _p(poses), _p(rec), _p(lig)))
        return rec, lig

    def energy_device(self, n, d_poses_ptr, d_energies_ptr, stream_ptr=None):
        """Device-resident poses/energies (raw pointers), asynchronous on `stream_ptr`."""
        _check(self.lib, self.lib.ld_score_batch_device(self.h, int(n), C.c_void_p(d_poses_ptr),
                                                        C.c_void_p(d_energies_ptr), C.c_void_p(stream_ptr or 0)))

    def set_rec_splits(self, splits):
        _check(self.lib, self.lib.ld_set_rec_splits(self.h, int(splits)))

    def set_path(self, path):
        """PATH_AUTO / PATH_GENERIC / PATH_RIGID (include/lightdock_b200.h: ld_set_path)."""
        _check(self.lib, self.lib.ld_set_path(self.h, int(path)))

    def path_info(self):
        return self.lib.ld_path_info(self.h).decode()

    def create_ms(self):
        """ld_create's time split: CUDA context, complex, receptor groups, cell lists (ms)."""
        out = (C.c_double * 4)()
        _check(self.lib, self.lib.ld_get_create_ms(self.h, out))
        return dict(zip(("context", "complex", "groups", "cells"), out))

    def set_profiling(self, on):
        _check(self.lib, self.lib.ld_set_profiling(self.h, int(bool(on))))

    def stats(self, slot=None):
        return handle_stats(self.lib, self.h, slot)


def set_option(key, value):
    """Process-wide tuning default for handles created afterwards (include/lightdock_b200.h: ld_set_option)."""
    lib = load_library()
    _check(lib, lib.ld_set_option(key.encode(), float(value)))


def handle_stats(lib, h, slot=None):
    s = BatchStats()
    if slot is None:
        _check(lib, lib.ld_get_stats(h, C.byref(s)))
    else:
        _check(lib, lib.ld_get_stats_slot(h, int(slot), C.byref(s)))
    return dict(n_poses=s.n_poses, pair_evals_bruteforce=s.pair_evals_bruteforce,
                kernel_launches=s.kernel_launches, rec_splits=s.rec_splits, device_ms=s.device_ms,
                transform_ms=s.transform_ms, pair_ms=s.pair_ms, finalize_ms=s.finalize_ms, path=s.path,
                pair_launches=s.pair_launches)


def probe_peaks(device=0):
    """On-box roofline denominators: (fp64 non-FMA TFLOP/s, fp32 non-FMA TFLOP/s, L2 gather Gloads/s)."""
    lib = load_library()
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    _check(lib, lib.ld_probe_peaks(int(device), C.addressof(a), C.addressof(b), C.addressof(c)))
    return a.value, b.value, c.value
