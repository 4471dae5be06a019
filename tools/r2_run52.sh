#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rigid_path.py -m gpu -x -q 2>&1 | tail -3
for r in 8 6 4 3; do echo "== LDB200_ROWS=$r"; LDB200_ROWS=$r timeout 600 python tools/ab_rigid.py 2>&1 | grep -E "80000|1ppe|2uuy|1czy"; done > gpurun_out/r2_run52_rows.log 2>&1
cat gpurun_out/r2_run52_rows.log
python - <<'PY'
import sys
for p in ("", "oracle", "lightdock-rust_b200", "tests", "tools"):
    sys.path.insert(0, p)
import oracle as O
from helpers import case, scorer_from_oracle
for name in ("1czy", "1ppe", "2uuy", "ab_icode", "1k4c"):
    cx, pos, _ = case(name, O.DFIRE)
    sc = scorer_from_oracle(cx)
    print(name, sc.path_info()[:150])
PY
