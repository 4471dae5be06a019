"""ctypes binding of liblightdock_host.so: the C++ host layer (model building, Score interface, GSO)
that the drop-in CLI `bin/lightdock-rust` is made of.  Product code; never imports oracle/."""
import ctypes as C
import os

import numpy as np

from . import PKG_DIR, LdError

HOST_LIB_PATH = os.path.join(PKG_DIR, "liblightdock_host.so")
CLI_PATH = os.path.join(PKG_DIR, "bin", "lightdock-rust")

_lib = None


def load_host_library():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise LdError(f"{HOST_LIB_PATH} not built: run `make -C lightdock-rust_b200`")
        lib = C.CDLL(HOST_LIB_PATH)
        lib.ldh_last_error.restype = C.c_char_p
        lib.ldh_open_case.restype = C.c_void_p
        lib.ldh_open_case.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        lib.ldh_close_case.argtypes = [C.c_void_p]
        lib.ldh_case_handle.restype = C.c_void_p
        lib.ldh_case_handle.argtypes = [C.c_void_p]
        lib.ldh_case_info.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        lib.ldh_case_model.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        lib.ldh_case_restraints.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
        lib.ldh_case_energy_batch.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]
        lib.ldh_case_energy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                        C.c_void_p]
        lib.ldh_case_gso.argtypes = [C.c_void_p, C.c_char_p, C.c_uint, C.c_char_p, C.c_void_p, C.c_void_p]
        lib.ldh_case_multi_gso.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ldh_case_device_gso.argtypes = lib.ldh_case_multi_gso.argtypes
        lib.ldh_rng_draws.restype = C.c_double
        lib.ldh_rng_draws.argtypes = [C.c_ulonglong, C.c_int, C.c_void_p]
        lib.ldh_slerp.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        lib.ldh_rotate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ldh_find_neighbors.argtypes = [C.c_int] + [C.c_void_p] * 5
        lib.ldh_shard_swarms.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                         C.c_void_p]
        lib.ldh_parse_f64.argtypes = [C.c_char_p, C.c_void_p]
        lib.ldh_save_swarm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint, C.c_char_p]
        lib.ldh_build_model.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p] + [C.c_void_p] * 10
        _lib = lib
    return _lib


def _err(lib):
    return LdError(lib.ldh_last_error().decode())


def rng_draws(seed, n):
    out = np.empty(n, dtype=np.float64)
    load_host_library().ldh_rng_draws(seed, n, out.ctypes.data)
    return out


def slerp(a, b, t):
    a, b, out = np.asarray(a, np.float64), np.asarray(b, np.float64), np.empty(4)
    load_host_library().ldh_slerp(a.ctypes.data, b.ctypes.data, t, out.ctypes.data)
    return out


def rotate(q, v):
    q, v, out = np.asarray(q, np.float64), np.asarray(v, np.float64), np.empty(3)
    load_host_library().ldh_rotate(q.ctypes.data, v.ctypes.data, out.ctypes.data)
    return out


def shard_swarms(rec_xyz, lig_xyz, centres, n_gpus):
    """Host-only: the cost-aware, deterministic swarm -> GPU map of host/sharding.hpp (what lightdock-rust-multi and
    bench.py use).  centres [n_swarms][3] = mean translation of each swarm.  Returns (gpu_of [n_swarms], cost)."""
    lib = load_host_library()
    rec = np.ascontiguousarray(rec_xyz, np.float64).reshape(-1, 3)
    lig = np.ascontiguousarray(lig_xyz, np.float64).reshape(-1, 3)
    cen = np.ascontiguousarray(centres, np.float64).reshape(-1, 3)
    cost = np.zeros(len(cen))
    gpu = np.zeros(len(cen), np.int32)
    if lib.ldh_shard_swarms(len(rec), rec.ctypes.data, len(lig), lig.ctypes.data, len(cen), cen.ctypes.data, int(n_gpus),
                            cost.ctypes.data, gpu.ctypes.data):
        raise _err(lib)
    return gpu, cost


def parse_f64(token):
    """Host-only: `token.parse::<f64>()` as the start-position reader applies it; None = ParseFloatError."""
    v = C.c_double()
    return v.value if load_host_library().ldh_parse_f64(token.encode(), C.addressof(v)) else None


def save_swarm(rows, n_rec_anm, n_lig_anm, step, out_dir):
    """Host-only: Swarm::save on a given state; rows [n][4 + pose_len] = luciferin, scoring, vision range,
    neighbour count, pose."""
    lib = load_host_library()
    rows = np.ascontiguousarray(rows, np.float64)
    if lib.ldh_save_swarm(rows.shape[0], n_rec_anm, n_lig_anm, rows.ctypes.data, step, out_dir.encode()):
        raise _err(lib)


def find_neighbors(xyz, luciferin, vision_range):
    """Host-only: the neighbour search of Swarm::movement_phase (src/swarm.rs:85-103) on a given swarm state."""
    lib = load_host_library()
    xyz = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
    n = xyz.shape[0]
    lum = np.ascontiguousarray(luciferin, np.float64)
    vr = np.ascontiguousarray(vision_range, np.float64)
    off = np.zeros(n + 1, np.int32)
    idx = np.zeros(max(1, n * (n - 1)), np.int32)
    k = lib.ldh_find_neighbors(n, xyz.ctypes.data, lum.ctypes.data, vr.ctypes.data, off.ctypes.data, idx.ctypes.data)
    if k < 0:
        raise _err(lib)
    return [idx[off[i]:off[i + 1]].tolist() for i in range(n)]


def build_model(pdb_path, method, active_restraints=()):
    """Host-only (no GPU): numeric model of one structure as the C++ layer types it."""
    lib = load_host_library()
    rst = "\n".join(active_restraints).encode()
    n = lib.ldh_build_model(pdb_path.encode(), method.encode(), rst, *([None] * 10))
    if n < 0:
        raise _err(lib)
    types = np.zeros(n, np.int32); coords = np.zeros((n, 3)); ele = np.zeros(n); ve = np.zeros(n); vr = np.zeros(n)
    nmem = C.c_int(); ngrp = C.c_int()
    mem = np.zeros(n, np.int32); off = np.zeros(n + 2, np.int32); idx = np.zeros(n + 1, np.int32)
    lib.ldh_build_model(pdb_path.encode(), method.encode(), rst, types.ctypes.data, coords.ctypes.data,
                        ele.ctypes.data, ve.ctypes.data, vr.ctypes.data, C.addressof(nmem), mem.ctypes.data,
                        C.addressof(ngrp), off.ctypes.data, idx.ctypes.data)
    g = ngrp.value
    return dict(n=n, dfire_type=types, coords=coords, ele=ele, vdw_e=ve, vdw_r=vr, membrane=mem[:nmem.value].copy(),
                rst_offsets=off[:g + 1].copy(), rst_atoms=idx[:off[g]].copy())


class Case:
    """setup.json + method -> device-resident scoring object, as `simulate` builds it."""

    def __init__(self, setup_json, method, anm_dir=None, device=0):
        self.lib = load_host_library()
        self.c = self.lib.ldh_open_case(setup_json.encode(), method.encode(), (anm_dir or "").encode(), device)
        if not self.c:
            raise _err(self.lib)
        v = [C.c_int() for _ in range(4)]
        seed = C.c_ulonglong()
        self.lib.ldh_case_info(self.c, *[C.addressof(x) for x in v], C.addressof(seed))
        self.n_rec, self.n_lig, self.pose_len, self.use_anm = (x.value for x in v)
        self.seed = seed.value

    def close(self):
        if getattr(self, "c", None):
            self.lib.ldh_close_case(self.c)
            self.c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ld_handle(self):
        return self.lib.ldh_case_handle(self.c)

    def path_info(self):
        """Which pair kernel the scoring object selected (include/lightdock_b200.h: ld_path_info)."""
        from . import load_library
        return load_library().ld_path_info(self.ld_handle()).decode()

    def energy_batch(self, poses):
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, self.pose_len)
        out = np.empty(poses.shape[0])
        if self.lib.ldh_case_energy_batch(self.c, poses.shape[0], poses.ctypes.data, out.ctypes.data):
            raise _err(self.lib)
        return out

    def energy(self, translation, quat, rec_nm=(), lig_nm=()):
        t, q = np.asarray(translation, np.float64), np.asarray(quat, np.float64)
        r, l = np.asarray(rec_nm, np.float64), np.asarray(lig_nm, np.float64)
        out = C.c_double()
        if self.lib.ldh_case_energy(self.c, t.ctypes.data, q.ctypes.data, r.ctypes.data, r.size, l.ctypes.data, l.size,
                                    C.addressof(out)):
            raise _err(self.lib)
        return out.value

    def gso(self, positions_file, steps, out_dir=None, n_glowworms=200):
        state = np.zeros((n_glowworms, 4 + self.pose_len))
        calls = C.c_ulonglong()
        if self.lib.ldh_case_gso(self.c, positions_file.encode(), steps, (out_dir or "").encode(), state.ctypes.data,
                                 C.addressof(calls)):
            raise _err(self.lib)
        return state, calls.value

    def device_gso(self, positions, seeds, steps, host_threads=1, out_dirs=None):
        """multi_gso with the whole GSO step on the device (DeviceGSO -> ld_gso_*): same arguments, same results up to
        the last-bit difference of CUDA's acos/sin inside slerp."""
        return self.multi_gso(positions, seeds, steps, host_threads, out_dirs, _entry="ldh_case_device_gso")

    def multi_gso(self, positions, seeds, steps, host_threads=1, out_dirs=None, _entry="ldh_case_multi_gso"):
        positions = np.ascontiguousarray(positions, np.float64)
        ns, ng, pl = positions.shape
        assert pl == self.pose_len
        seeds = np.ascontiguousarray(seeds, np.uint64)
        state = np.zeros((ns, ng, 4 + pl))
        calls = C.c_ulonglong()
        dirs = None
        if out_dirs is not None:
            dirs = (C.c_char_p * ns)(*[d.encode() for d in out_dirs])
        if getattr(self.lib, _entry)(self.c, ns, ng, positions.ctypes.data, seeds.ctypes.data, steps, host_threads,
                                     dirs, state.ctypes.data, C.addressof(calls)):
            raise _err(self.lib)
        return state, calls.value
