#!/bin/bash
# round-2 GPU run 42: validation of the kernels as shipped: compute-sanitizer (memcheck, racecheck) on small invocations of
# every kernel, then the randomized parity campaign
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/r2f_sanitizer_memcheck.log 2>&1
tail -3 gpurun_out/r2f_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/r2f_sanitizer_racecheck.log 2>&1
tail -3 gpurun_out/r2f_sanitizer_racecheck.log
POSES=400000 FLEX_POSES=100000 timeout 2400 python tools/parity_campaign.py > gpurun_out/r2f_parity_campaign.log 2>&1
tail -12 gpurun_out/r2f_parity_campaign.log
