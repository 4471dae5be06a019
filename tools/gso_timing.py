"""Device-resident GSO against the host loop (run under gpurun): 400 synthetic 1k4c swarms x 20 steps, three repetitions each,
then the 1azp CLI with LIGHTDOCK_GSO=device five times (gso_ms of its timing line)."""
import os, re, shutil, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("", "lightdock-rust_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
from ldb200 import host, workload
dc_dir, _ = workload.ensure_dcparams_dir("/tmp")
os.environ["LIGHTDOCK_DATA"] = dc_dir
c = host.Case(os.path.join(workload.GOLDEN_1K4C, "setup.json"), "dfire")
pos = workload.synthetic_1k4c_swarms(400, 200)
seeds = np.full(400, 324324, dtype=np.uint64)
c.multi_gso(pos, seeds, 1, host_threads=16); c.device_gso(pos, seeds, 1, host_threads=16)
for rep in range(6):
    t = time.perf_counter(); _, calls = c.multi_gso(pos, seeds, 20, host_threads=16); th = time.perf_counter() - t
    t = time.perf_counter(); _, calls_d = c.device_gso(pos, seeds, 20, host_threads=16); td = time.perf_counter() - t
    print(f"rep {rep}: host loop {th * 1e3:.1f} ms ({calls / th / 1e6:.2f} M calls/s), device loop {td * 1e3:.1f} ms ({calls_d / td / 1e6:.2f} M calls/s)", flush=True)
g = os.path.join(ROOT, "tests", "golden", "1azp")
for mode in ("host", "device"):
    ms = []
    for rep in range(5):
        with tempfile.TemporaryDirectory() as tmp:
            for f in ("rec_nm.npy", "lig_nm.npy"):
                shutil.copy(os.path.join(g, f), os.path.join(tmp, f))
            env = dict(os.environ, LDB200_TIMING="1")
            if mode == "device":
                env["LIGHTDOCK_GSO"] = "device"
            t = time.perf_counter()
            r = subprocess.run([host.CLI_PATH, os.path.join(g, "setup.json"), os.path.join(g, "initial_positions_0.dat"), "100", "dna"],
                               cwd=tmp, env=env, capture_output=True, text=True)
            wall = time.perf_counter() - t
            m = re.search(r"gso_ms=([0-9.]+)", r.stderr)
            ms.append((float(m.group(1)) if m else None, round(wall, 3)))
    print("1azp CLI", mode, "(gso_ms, wall_s):", ms, flush=True)
