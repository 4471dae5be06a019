#!/bin/bash
# round-2 GPU run 13: compute-sanitizer on every kernel incl. the round-2 ones; FLEX rows-per-group A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool python tools/sanitize_small.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|device gso|flex lists" gpurun_out/r2_sanitizer_$tool.log | tail -4
done
for r in 4 3 2 1; do echo "== LDB200_ROWS=$r"; LDB200_ROWS=$r timeout 600 python tools/ab_rigid.py 2>&1 | tail -3; done > gpurun_out/r2_run13_rows.log 2>&1
cat gpurun_out/r2_run13_rows.log
