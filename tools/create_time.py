import os, sys, time
ROOT = os.getcwd()
for p in ("", "oracle", "lightdock-rust_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import oracle as O
from helpers import case, scorer_from_oracle
for name in ("1ppe", "1k4c"):
    cx, pos, _ = case(name, O.DFIRE)
    sc = scorer_from_oracle(cx); sc.close()
    t = time.time(); sc = scorer_from_oracle(cx); dt = time.time() - t
    print(name, f"ld_create {dt*1e3:.1f} ms;", sc.path_info()[:200])
