"""DNA / pyDock kernel timing on 1azp (run under gpurun): large-batch throughput and per-kernel times."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("", "oracle", "lightdock-rust_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import oracle as O
from helpers import case, scorer_from_oracle, assert_parity
cx, pos, _ = case("1azp", O.DNA)
sc = scorer_from_oracle(cx)
e_gpu, d_gpu = sc.energy_detail(pos[:32]); e_ref, d_ref = cx.energy(pos[:32], detail=True)
print("parity rel", assert_parity(e_gpu, d_gpu, e_ref, d_ref, cx.method))
rng = np.random.default_rng(1)
big = np.tile(pos, (100, 1)); big[:, :3] += rng.normal(0, 1.0, size=(len(big), 3))
sc.set_profiling(True)
sc.energy(big)
for _ in range(2):
    t = time.perf_counter(); sc.energy(big); dt = time.perf_counter() - t
    st = sc.stats()
    print(f"{len(big)} poses {dt*1e3:.2f} ms  transform {st['transform_ms']:.2f} pair {st['pair_ms']:.2f} finalize {st['finalize_ms']:.2f} -> {len(big)/dt/1e6:.3f} M poses/s, "
          f"{len(big)*cx.rec.n*cx.lig.n/st['pair_ms']*1e3/1e12:.3f} T pair-evals/s in the pair kernel")
_, d = sc.energy_detail(big[:2000])
n = cx.rec.n * cx.lig.n
print("fractions: elec", d["n_in_cutoff"].mean() / n, "vdw", d["n_in_cutoff2"].mean() / n, "tested", d["n_pairs_tested"].mean() / n)
