"""bench.py prints ONE JSON line with the keys the driver and the judge read (task contract ④); run here on a
reduced workload so the schema cannot rot unnoticed."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_b200_arm_line():
    d = _run(["--swarms", "24", "--steps", "3", "--warmup", "3", "--cpu-seconds", "1", "--gso-steps", "3",
              "--no-single-swarm-runs", "--config-cpu-seconds", "0.3"])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "glowworm_poses_scored_per_s" and d["unit"] == "poses/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["vs_baseline"] is None
    assert "workload" in d["config"] and d["config"]["poses_per_step"] == 24 * 200
    assert d["value"] > 1e5 and d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["unit"] == "poses/s" and e["value"] > 1e5 and e["h2d_bytes_per_step"] == 24 * 200 * 7 * 8
    assert e["d2h_bytes_per_step"] == 24 * 200 * 8
    rf = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel"):
        assert k in rf, k
    assert rf["kernel"] == "dfire_rigid_kernel" and rf["bound"] == "issue" and rf["unit"] == "Gwarp-inst/s"
    # the fraction is printed only with ncu counters whose source hash matches this build; otherwise bench.py says why
    if rf["frac"] is not None:
        assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and 0.0 < rf["frac"] < 1.0
        assert 0.0 < rf["issue"]["smem_wavefront_frac"] < 1.0
    else:
        assert "counts_refused" in rf
    assert rf["bruteforce_equivalent"]["ratio"] > 1.0 and 0 < rf["hbm"]["frac"] < 0.05
    # BASELINE configs 0-3 (N=1): e2e poses/s, dominant kernel, roofline, CPU port on the same poses
    cf = d["configs"]
    assert set(cf) == {"1czy", "1ppe", "2uuy", "1azp"}
    for name, v in cf.items():
        assert v["e2e"]["value"] > 1e5 and v["poses"] == 20000 and v["cpu_baseline"]["value"] > 0
        assert v["gpu_vs_oracle_max_rel_err_8_poses"] < 1e-6, name
        assert v["roofline"]["kernel_ms"] > 0
        assert v["roofline"]["kernel"] == ("dna_pair_kernel" if name == "1azp" else "dfire_rigid_kernel"), (name, v["path_info"])
    assert cf["1azp"]["roofline"]["bound"] == "fp64" and 0.2 < cf["1azp"]["roofline"]["frac"] < 1.5
    g = d["gso_run"]
    assert g["pair_tests_per_pose_start"] > 0 and g["pair_tests_per_pose_after_gso"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    assert c["gpu_vs_oracle_max_rel_err_8_poses"] < 1e-6
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "2", "--warmup", "1"])
    assert d["impl"] == "reference" and d["metric"] == "glowworm_poses_scored_per_s" and d["value"] > 0
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"] == {"value": d["value"], "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
