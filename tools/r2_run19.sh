#!/bin/bash
# round-2 GPU run 19: compact tiles (capacity-constrained k-means on top of the bisection order): tests + A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_run19_pytest.log
tail -8 gpurun_out/r2_run19_pytest.log
for c in 0 1; do
  export LDB200_COMPACT_TILES=$c
  echo "== LDB200_COMPACT_TILES=$c"
  timeout 600 python tools/ab_rigid.py 2>&1 | tail -6
  timeout 300 python tools/dna_timing.py 2>&1 | grep -E "poses|fractions" | tail -2
  timeout 600 python tools/config_timings.py 2>&1 | tail -5
done > gpurun_out/r2_run19_ab.log 2>&1
cat gpurun_out/r2_run19_ab.log
