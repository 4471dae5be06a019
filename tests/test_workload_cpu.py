"""CPU tests of the benchmark workload generator and the swarm sharding used for N > 1 GPUs, including a
world_size-2 gloo run of the aggregation logic bench.py uses (max-over-ranks time, summed pose counts)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

import oracle as O
from helpers import ROOT
from ldb200 import workload


def test_synthetic_table_is_the_same_in_product_and_oracle():
    a, b = workload.synthetic_dcparams(), O.synthetic_dcparams()
    assert a.size == 169 * 169 * 20 and np.array_equal(a, b)
    assert (a[:2] == 10.0).all() and a[-1] == 0.0


def test_table_text_round_trip(tmp_path):
    t = workload.synthetic_dcparams()
    workload.write_dcparams(tmp_path / "DCparams", t)
    assert np.array_equal(O.load_dcparams(str(tmp_path / "DCparams")), t)


def test_synthetic_swarms_shape_and_statistics():
    s = workload.synthetic_1k4c_swarms(8, 200)
    assert s.shape == (8, 200, 7)
    assert np.allclose(np.linalg.norm(s[..., 3:7], axis=-1), 1.0)
    centres = np.loadtxt(os.path.join(workload.DATA_DIR, "synthetic_1k4c_swarm_centres.tsv"), comments="#")
    assert centres.shape == (400, 3)
    assert (np.linalg.norm(s[:, :, :3] - centres[:8, None, :], axis=-1) <= 10.0 + 1e-9).all()
    again = workload.synthetic_1k4c_swarms(8, 200)
    assert np.array_equal(s, again)  # seeded


def test_synthetic_swarms_hit_the_measured_in_cutoff_range():
    """0.4-2 % of the atom pairs of a pose are inside DFIRE's 15 A cut-off."""
    from helpers import case
    cx, _, _ = case("1k4c", O.DFIRE)
    s = workload.synthetic_1k4c_swarms(400, 2)
    poses = s[::40, 0, :]
    _, d = cx.energy(poses, detail=True)
    frac = d["n_in_cutoff"] / (cx.rec.n * cx.lig.n)
    assert 0.002 < frac.min() and frac.max() < 0.035 and 0.006 < frac.mean() < 0.02


def test_sharding_is_a_partition():
    for world in (1, 2, 4, 8):
        owned = [workload.shard_swarms(400, r, world) for r in range(world)]
        flat = sorted(x for o in owned for x in o)
        assert flat == list(range(400))
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_gloo_aggregation(tmp_path):
    """N > 1 host logic on CPU: each rank owns the swarms the cost-aware map gives it, results are combined the way
    bench.py does (sum of poses, MAX of time), and the union of the shards equals the single-rank pose set."""
    script = tmp_path / "rank.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, json
        import numpy as np, torch, torch.distributed as dist
        sys.path.insert(0, {os.path.join(ROOT, 'lightdock-rust_b200')!r})
        from ldb200 import workload
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        dist.init_process_group("gloo")
        g = workload.GOLDEN_1K4C
        rec = workload.read_pdb_coords(g + "/lightdock_receptor_membrane.pdb")
        lig = workload.read_pdb_coords(g + "/lightdock_ligand.pdb")
        all_poses = workload.synthetic_1k4c_swarms(12, 5)
        mine = workload.shard_swarms_cost_aware(all_poses, rec, lig, rank, world)   # what bench.py does
        poses = all_poses[mine].reshape(-1, 7)
        t = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)      # pretend per-rank elapsed ms
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n = torch.tensor([poses.shape[0]], dtype=torch.int64)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        chk = torch.tensor([float(poses.sum())], dtype=torch.float64)
        dist.all_reduce(chk, op=dist.ReduceOp.SUM)
        if rank == 0:
            print(json.dumps(dict(max_ms=t.item(), n=int(n.item()), checksum=chk.item())))
        dist.destroy_process_group()
    """))
    port = _free_port()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    full = workload.synthetic_1k4c_swarms(12, 5).reshape(-1, 7)
    assert out["n"] == 60 and out["max_ms"] == 15.0
    assert abs(out["checksum"] - full.sum()) < 1e-6


def test_cost_aware_sharding_is_a_balanced_deterministic_partition():
    """host/sharding.hpp: every swarm on exactly one GPU, the same map on every call, and a flatter predicted load
    than the round-1 `s mod G` map on the bench workload (whose static map left the slowest of 8 ranks 10 % above the
    mean)."""
    from ldb200 import host
    g = workload.GOLDEN_1K4C
    rec = workload.read_pdb_coords(g + "/lightdock_receptor_membrane.pdb")
    lig = workload.read_pdb_coords(g + "/lightdock_ligand.pdb")
    poses = workload.synthetic_1k4c_swarms(400, 8)
    centres = poses[:, :, :3].mean(axis=1)
    for G in (2, 4, 8):
        gpu, cost = host.shard_swarms(rec, lig, centres, G)
        gpu2, _ = host.shard_swarms(rec, lig, centres, G)
        assert np.array_equal(gpu, gpu2) and set(gpu.tolist()) == set(range(G))
        owned = [workload.shard_swarms_cost_aware(poses, rec, lig, r, G) for r in range(G)]
        assert sorted(sum(owned, [])) == list(range(400))
        lpt = np.array([cost[gpu == r].sum() for r in range(G)])
        mod = np.array([cost[r::G].sum() for r in range(G)])
        assert lpt.max() / lpt.mean() < 1.005 and lpt.max() / lpt.mean() <= mod.max() / mod.mean()
    assert (cost > 0).all() and cost.max() / cost.min() > 3, "the bench swarms do differ in cost"
    # a swarm far from the receptor costs (almost) nothing, one on top of it the most
    _, c = host.shard_swarms(rec, lig, np.array([[500.0, 0, 0], rec.mean(axis=0)]), 2)
    assert c[0] < 2 and c[1] > 1e5


def test_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` needs no GPU and prints the contract line."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--swarms", "4", "--glowworms", "50"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "poses/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
