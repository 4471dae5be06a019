#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_run38_pytest.log
cat gpurun_out/r2_run38_pytest.log
(echo "== auto"; timeout 600 python tools/ab_rigid.py 2>&1 | grep -E "poses|parity"
echo "== LDB200_CELL=0.5"; LDB200_CELL=0.5 timeout 600 python tools/ab_rigid.py 2>&1 | grep -E "1ppe|1czy|2uuy") > gpurun_out/r2_run38_cell.log 2>&1
cat gpurun_out/r2_run38_cell.log
