"""Quick GPU bring-up script (not a test): parity spot-check + first timing. Run under gpurun."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("", "oracle", "lightdock-rust_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import oracle as O
from helpers import case, scorer_from_oracle

for name, method in (("1azp", O.DNA), ("2uuy", O.DFIRE), ("1k4c", O.DFIRE)):
    cx, pos, _ = case(name, method)
    sc = scorer_from_oracle(cx)
    n = 16
    e_gpu, d = sc.energy_detail(pos[:n])
    e_ref, dr = cx.energy(pos[:n], detail=True)
    print(name, "max rel err", np.max(np.abs(e_gpu - e_ref) / np.abs(e_ref)), "in_cutoff equal",
          np.array_equal(d["n_in_cutoff"], dr["n_in_cutoff"]), "hist equal", np.array_equal(d["bin_hist"], dr["bin_hist"]),
          "iface equal", np.array_equal(d["iface_rec"], dr["iface_rec"]) and np.array_equal(d["iface_lig"], dr["iface_lig"]))
    print("   in-cutoff fraction", dr["n_in_cutoff"].mean() / (cx.rec.n * cx.lig.n))
    big = np.tile(pos, (max(1, 8000 // len(pos)), 1))
    sc.energy(big[:200])
    for reps in range(2):
        t = time.time(); e = sc.energy(big); dt = time.time() - t
        st = sc.stats()
        print(f"   {len(big)} poses: wall {dt*1e3:.1f} ms, device {st['device_ms']:.1f} ms -> {len(big)/st['device_ms']*1e3:.0f} poses/s, "
              f"{st['pair_evals_bruteforce']/st['device_ms']*1e3:.3e} pair-evals/s, splits {st['rec_splits']}")
