#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for mb in 64 32 16 8 128; do echo "== LDB200_FLEX_CHUNK_MB=$mb"; LDB200_FLEX_CHUNK_MB=$mb timeout 600 python tools/config_timings.py 2>&1 | grep -E "1czy|2uuy"; done > gpurun_out/r2_run26_flexchunk.log 2>&1
cat gpurun_out/r2_run26_flexchunk.log
