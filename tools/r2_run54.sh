#!/bin/bash
# round-2 GPU run 54: final build: GPU suite, sanitizer, parity campaign, counters + launch list + bench + captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_run54_pytest.log; cat gpurun_out/r2_run54_pytest.log
timeout 1500 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/r2f_sanitizer_memcheck.log 2>&1
tail -1 gpurun_out/r2f_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/r2f_sanitizer_racecheck.log 2>&1
tail -1 gpurun_out/r2f_sanitizer_racecheck.log
POSES=400000 FLEX_POSES=100000 timeout 2400 python tools/parity_campaign.py > gpurun_out/r2f_parity_campaign.log 2>&1
tail -6 gpurun_out/r2f_parity_campaign.log
bash tools/r2_run39.sh
