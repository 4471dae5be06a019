"""GPU: the device-resident GSO (include/lightdock_b200.h ld_gso_*, host/gso.hpp DeviceGSO) against
the host loop and the oracle.  The step kernel restates src/swarm.rs:66-126 and src/glowworm.rs:61-190 in the
reference's f64 operation order and evaluates each swarm's ChaCha20 stream on the device, so every discrete field
(neighbour counts, vision ranges, number of energy evaluations) must equal the host's exactly; poses, scores and
luciferin agree to the north-star tolerance (the one rounding difference is CUDA's acos/sin inside slerp)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle as O
from helpers import ENERGY_RTOL, GOLDEN, case
from test_gpu_trajectory import STEPS, compare_gso_files, oracle_threads

pytestmark = pytest.mark.gpu

POSE_ATOL = 1e-9  # absolute, on translations of tens of A and unit quaternions


def assert_same_state(dev, ref):
    np.testing.assert_array_equal(dev[..., 2], ref[..., 2], err_msg="neighbour counts")
    np.testing.assert_array_equal(dev[..., 3], ref[..., 3], err_msg="vision range")
    assert np.abs(dev[..., 4:] - ref[..., 4:]).max() <= POSE_ATOL, "poses"
    assert (np.abs(dev[..., 1] - ref[..., 1]) <= ENERGY_RTOL * np.abs(ref[..., 1])).all(), "scoring"
    assert (np.abs(dev[..., 0] - ref[..., 0]) <= ENERGY_RTOL * np.abs(ref[..., 0]) + 1e-12).all(), "luciferin"


def start_positions(g):
    return np.array([[float(x) for x in l.split(" ")] for l in
                     open(os.path.join(g, "initial_positions_0.dat")).read().splitlines()])


def test_device_gso_reproduces_1azp_golden_trajectory_through_the_cli(tmp_path):
    """example/1azp (DNA + ANM 10/10 + restraints, seed 324324, 100 steps) with LIGHTDOCK_GSO=device: the drop-in CLI
    writes the same 11 files as the reference, to print precision."""
    from ldb200 import host
    g = os.path.join(GOLDEN, "1azp")
    for f in ("rec_nm.npy", "lig_nm.npy"):
        shutil.copy(os.path.join(g, f), tmp_path / f)
    env = dict(os.environ, LIGHTDOCK_GSO="device")
    r = subprocess.run([host.CLI_PATH, os.path.join(g, "setup.json"), os.path.join(g, "initial_positions_0.dat"),
                        "100", "dna"], cwd=tmp_path, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr
    assert r.stdout.splitlines()[-1] == "Starting optimization (100 steps)"
    assert sorted(os.listdir(tmp_path / "swarm_0")) == sorted(f"gso_{s}.out" for s in STEPS)
    for s in STEPS:
        compare_gso_files(str(tmp_path / "swarm_0" / f"gso_{s}.out"), os.path.join(g, "swarm_0", f"gso_{s}.out"))


def test_device_gso_equals_host_gso_dna_with_anm():
    """Three swarms of 1azp (the shipped one, a jittered copy, another seed), 100 steps: device loop == host loop."""
    from ldb200 import host
    g = os.path.join(GOLDEN, "1azp")
    c = host.Case(os.path.join(g, "setup.json"), "dna", anm_dir=g)
    pos = start_positions(g)
    pos2 = pos.copy()
    pos2[:, :3] += np.random.default_rng(5).normal(0, 0.3, size=(pos.shape[0], 3))
    start, seeds = np.stack([pos, pos2, pos]), [c.seed, c.seed, 99]
    ref, calls_ref = c.multi_gso(start, seeds, 100, host_threads=3)
    dev, calls_dev = c.device_gso(start, seeds, 100, host_threads=3)
    assert calls_dev == calls_ref
    assert_same_state(dev, ref)
    assert not np.array_equal(dev[0], dev[2]), "another seed is another trajectory"


@pytest.mark.timeout(1500)
@pytest.mark.parametrize("name", ["1ppe", "1k4c", "2uuy"])
def test_device_gso_matches_oracle_gso_dfire(name, tmp_path, monkeypatch):
    """DFIRE, 100 steps, against the ORACLE's GSO (not the product's host loop): 1ppe and 1k4c on the rigid ligand-frame
    kernel, 2uuy (ANM on both partners) on its FLEX instance.  Every saved step through the gso files, the final state
    at full precision."""
    from ldb200 import host
    cx, pos, seed = case(name, O.DFIRE)
    g = os.path.join(GOLDEN, name)
    O.write_dcparams(str(tmp_path / "DCparams"), cx.potential)
    monkeypatch.setenv("LIGHTDOCK_DATA", str(tmp_path))
    c = host.Case(os.path.join(g, "setup.json"), "dfire", anm_dir=g)
    assert c.path_info().startswith("rigid path on"), c.path_info()
    os.makedirs(tmp_path / "gpu"); os.makedirs(tmp_path / "cpu")
    state, calls = c.device_gso(pos[None], [seed], 100, out_dirs=[str(tmp_path / "gpu")])
    final, tr, ocalls = cx.gso_run(pos, seed, 100, out_dir=str(tmp_path / "cpu"), trace=True, threads=oracle_threads())
    for s in STEPS:
        compare_gso_files(str(tmp_path / "gpu" / f"gso_{s}.out"), str(tmp_path / "cpu" / f"gso_{s}.out"))
    last = tr[-1]
    assert calls == ocalls
    np.testing.assert_array_equal(state[0][:, 2], last[:, 2])
    np.testing.assert_array_equal(state[0][:, 3], last[:, 3])
    assert np.abs(state[0][:, 4:] - last[:, 5:]).max() <= POSE_ATOL
    assert (np.abs(state[0][:, 1] - last[:, 1]) <= ENERGY_RTOL * np.abs(last[:, 1])).all()


def test_device_gso_many_swarms_and_chunked_runs(monkeypatch, tmp_path):
    """48 synthetic 1k4c swarms, 25 steps: the device loop equals the host MultiGSO on every swarm; running the steps in
    several ld_gso_run calls (1 + 9 + 10 + 5, what DeviceGSO does around the save points) is what one call gives; and
    the rows-beyond-the-live-count skipping of the scoring kernels leaves no trace (moved-only rescoring)."""
    import ctypes as C
    import ldb200
    from ldb200 import host, workload
    dc_dir, _ = workload.ensure_dcparams_dir(str(tmp_path))
    monkeypatch.setenv("LIGHTDOCK_DATA", dc_dir)
    c = host.Case(os.path.join(workload.GOLDEN_1K4C, "setup.json"), "dfire")
    pos = workload.synthetic_1k4c_swarms(48, 200)
    seeds = np.arange(48, dtype=np.uint64) + 324324
    ref, calls_ref = c.multi_gso(pos, seeds, 25, host_threads=8)
    dev, calls_dev = c.device_gso(pos, seeds, 25, host_threads=8)
    assert calls_dev == calls_ref
    assert calls_ref < 48 * 200 * 25, "only the glowworms that moved are rescored"
    assert_same_state(dev, ref)
    # the raw C ABI, one call of 25 steps
    lib = ldb200.load_library()
    g = C.c_void_p()
    flat = np.ascontiguousarray(pos, np.float64)
    assert lib.ld_gso_create(C.c_void_p(c.ld_handle()), 48, 200, flat.ctypes.data, seeds.ctypes.data, C.byref(g)) == 0, \
        lib.ld_last_error()
    assert lib.ld_gso_run(g, 25) == 0, lib.ld_last_error()
    assert lib.ld_gso_steps(g) == 25 and lib.ld_gso_energy_calls(g) == calls_ref
    poses = np.empty((48, 200, 7)); lum = np.empty((48, 200)); vis = np.empty((48, 200)); sc = np.empty((48, 200))
    nn = np.empty((48, 200), np.int32); failed = np.empty(48, np.int32)
    assert lib.ld_gso_state(g, poses.ctypes.data, lum.ctypes.data, vis.ctypes.data, sc.ctypes.data, nn.ctypes.data,
                            failed.ctypes.data) == 0
    lib.ld_gso_destroy(g)
    assert not failed.any()
    assert np.array_equal(poses, dev[..., 4:]) and np.array_equal(lum, dev[..., 0]) and np.array_equal(sc, dev[..., 1])
    assert np.array_equal(nn, dev[..., 2].astype(np.int32)) and np.array_equal(vis, dev[..., 3])


def test_device_gso_rng_stream_is_the_reference_stream():
    """One step of a swarm in which every glowworm sees every brighter one (huge vision range is not settable, so use
    the known-answer route instead): the device's draws are StdRng::seed_from_u64(seed).gen::<f64>() -- checked through
    the host mirror's stream (pinned to src/qt.rs:451-462 in test_host_cpu) by comparing a 30-step trajectory, whose
    roulette choices consume 30 x 200 draws, for two seeds that differ only in the stream."""
    from ldb200 import host
    g = os.path.join(GOLDEN, "1azp")
    c = host.Case(os.path.join(g, "setup.json"), "dna", anm_dir=g)
    pos = start_positions(g)
    for seed in (1, 2**63 + 12345):
        ref, _ = c.multi_gso(pos[None], [seed], 30)
        dev, _ = c.device_gso(pos[None], [seed], 30)
        assert_same_state(dev, ref)


def _raw_gso(c, pos, seeds, step_chunks):
    """The raw C ABI: ld_gso_create, one ld_gso_run per entry of step_chunks, ld_gso_state."""
    import ctypes as C
    import ldb200
    lib = ldb200.load_library()
    pos = np.ascontiguousarray(pos, np.float64)
    S, n, pl = pos.shape
    seeds = np.ascontiguousarray(seeds, np.uint64)
    g = C.c_void_p()
    rc = lib.ld_gso_create(C.c_void_p(c.ld_handle()), S, n, pos.ctypes.data, seeds.ctypes.data, C.byref(g))
    if rc != 0:
        return rc, lib.ld_last_error().decode()
    try:
        for k in step_chunks:
            assert lib.ld_gso_run(g, k) == 0, lib.ld_last_error()
        out = dict(poses=np.empty((S, n, pl)), lum=np.empty((S, n)), vis=np.empty((S, n)), sc=np.empty((S, n)),
                   nn=np.empty((S, n), np.int32), failed=np.empty(S, np.int32))
        assert lib.ld_gso_state(g, *[out[k].ctypes.data for k in ("poses", "lum", "vis", "sc", "nn", "failed")]) == 0
        out["steps"] = lib.ld_gso_steps(g)
        out["calls"] = lib.ld_gso_energy_calls(g)
    finally:
        lib.ld_gso_destroy(g)
    return 0, out


def test_device_gso_edge_shapes():
    """Swarm sizes that are not a multiple of the warp size, a swarm of one glowworm (never moves: one energy call),
    zero steps (start state, Glowworm::new values), chunked runs == one run, and the glowworm limit."""
    from ldb200 import host
    g = os.path.join(GOLDEN, "1azp")
    c = host.Case(os.path.join(g, "setup.json"), "dna", anm_dir=g)
    pos = start_positions(g)
    for n in (37, 1, 64):
        sub = np.stack([pos[:n], pos[50:50 + n]])
        ref, calls_ref = c.multi_gso(sub, [c.seed, 7], 15)
        dev, calls_dev = c.device_gso(sub, [c.seed, 7], 15)
        assert calls_dev == calls_ref
        assert_same_state(dev, ref)
        if n == 1:
            assert calls_dev == 2 and np.array_equal(dev[..., 4:], sub), "a lone glowworm is scored once and never moves"
    rc, st = _raw_gso(c, pos[None], [c.seed], [0])
    assert rc == 0 and st["steps"] == 0 and st["calls"] == 0
    assert np.array_equal(st["poses"][0], pos) and (st["lum"] == 5.0).all() and (st["vis"] == 0.2).all()
    assert (st["sc"] == 0.0).all() and not st["nn"].any() and not st["failed"].any()
    rc, one = _raw_gso(c, pos[None], [c.seed], [7])
    rc2, chunks = _raw_gso(c, pos[None], [c.seed], [3, 0, 4])
    assert rc == 0 and rc2 == 0 and one["steps"] == chunks["steps"] == 7 and one["calls"] == chunks["calls"]
    for k in ("poses", "lum", "vis", "sc", "nn"):
        assert np.array_equal(one[k], chunks[k]), k
    big = np.tile(pos[:1], (1025, 1))[None]
    rc, msg = _raw_gso(c, big, [1], [])
    assert rc == -4 and "LD_GSO_MAX_GLOWWORMS" in msg  # LD_ELIMIT
    rc, ok = _raw_gso(c, np.tile(pos[:1], (1024, 1))[None] + np.linspace(0, 1, 1024)[None, :, None], [1], [2])
    assert rc == 0 and ok["steps"] == 2 and not ok["failed"].any()


def test_device_gso_interleaves_with_ordinary_scoring_calls():
    """Between two ld_gso_run calls the handle's other entry points stay usable (they share slot 0's buffers and stream)
    and do not disturb the optimisation."""
    from ldb200 import host
    g = os.path.join(GOLDEN, "1azp")
    c = host.Case(os.path.join(g, "setup.json"), "dna", anm_dir=g)
    pos = start_positions(g)
    import ctypes as C
    import ldb200
    lib = ldb200.load_library()
    rc, ref = _raw_gso(c, pos[None], [c.seed], [12])
    gh = C.c_void_p()
    seeds = np.array([c.seed], np.uint64)
    flat = np.ascontiguousarray(pos[None], np.float64)
    assert lib.ld_gso_create(C.c_void_p(c.ld_handle()), 1, len(pos), flat.ctypes.data, seeds.ctypes.data, C.byref(gh)) == 0
    e0 = c.energy_batch(pos[:50])
    assert lib.ld_gso_run(gh, 5) == 0
    assert np.array_equal(c.energy_batch(pos[:50]), e0)
    assert lib.ld_gso_run(gh, 7) == 0
    poses = np.empty((1, len(pos), pos.shape[1]))
    assert lib.ld_gso_state(gh, poses.ctypes.data, None, None, None, None, None) == 0
    lib.ld_gso_destroy(gh)
    assert np.array_equal(poses, ref["poses"])


def test_multi_cli_with_device_gso(tmp_path):
    """lightdock-rust-multi with LIGHTDOCK_GSO=device: three swarms of 1azp, 12 steps, against the same run with the host
    loop -- same files, same discrete columns, poses / scores to print precision."""
    from ldb200 import host
    g = os.path.join(GOLDEN, "1azp")
    multi_cli = os.path.join(os.path.dirname(host.CLI_PATH), "lightdock-rust-multi")
    base = open(os.path.join(g, "initial_positions_0.dat")).read().splitlines()
    rng = np.random.default_rng(3)
    texts = {}
    for k in (0, 3, 11):
        rows = [[float(x) for x in l.split(" ")] for l in base]
        if k:
            for r in rows:
                r[0] += rng.normal(0, 0.5); r[1] += rng.normal(0, 0.5); r[2] += rng.normal(0, 0.5)
        texts[k] = "\n".join(" ".join(repr(v) for v in r) for r in rows) + "\n"
    files = [f"init/initial_positions_{k}.dat" for k in texts]
    for mode in ("host", "device"):
        d = tmp_path / mode
        os.makedirs(d / "init")
        for f in ("rec_nm.npy", "lig_nm.npy"):
            shutil.copy(os.path.join(g, f), d / f)
        for k, text in texts.items():
            (d / "init" / f"initial_positions_{k}.dat").write_text(text)
        env = dict(os.environ)
        if mode == "device":
            env["LIGHTDOCK_GSO"] = "device"
        r = subprocess.run([multi_cli, os.path.join(g, "setup.json"), "12", "dna"] + files, cwd=d, capture_output=True,
                           text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stderr
        assert "3 swarms, 12 steps" in r.stdout and "Done:" in r.stdout
    for k in texts:
        for step in (1, 10):
            compare_gso_files(str(tmp_path / "device" / f"swarm_{k}" / f"gso_{step}.out"),
                              str(tmp_path / "host" / f"swarm_{k}" / f"gso_{step}.out"))
        assert sorted(os.listdir(tmp_path / "device" / f"swarm_{k}")) == ["gso_1.out", "gso_10.out"]
