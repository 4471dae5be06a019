#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rigid_path.py -m gpu -x -q 2>&1 | tail -3
for w in 12 16 18 20; do echo "== LDB200_FLEX_MIN_WARPS=$w"; LDB200_FLEX_MIN_WARPS=$w timeout 600 python tools/ab_rigid.py 2>&1 | grep -E "80000|1ppe|2uuy|1czy"
LDB200_FLEX_MIN_WARPS=$w python - <<'PY'
import sys
for p in ("", "oracle", "lightdock-rust_b200", "tests", "tools"):
    sys.path.insert(0, p)
import numpy as np
import oracle as O
from helpers import case, scorer_from_oracle
from count_target import config_poses
for name in ("1czy", "2uuy", "ab_icode"):
    cx, pos, _ = case(name, O.DFIRE)
    sc = scorer_from_oracle(cx)
    info = sc.path_info()
    poses = np.ascontiguousarray(np.tile(pos, (100, 1)))
    rng = np.random.default_rng(1); poses[:, :3] += rng.normal(0, 1.0, size=(len(poses), 3))
    sc.set_profiling(True); sc.energy(poses); sc.energy(poses)
    ts = []
    for _ in range(3):
        sc.energy(poses); ts.append(sc.stats()["pair_ms"])
    print(name, "pair ms", np.round(ts, 3), info[info.index("):") + 2: info.index("cell")], info[info.index("smem"):])
PY
done > gpurun_out/r2_run53_flexrows.log 2>&1
cat gpurun_out/r2_run53_flexrows.log
