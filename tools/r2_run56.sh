#!/bin/bash
# round-2 GPU run 56: DNA kernel at 6 x 192 threads: GPU suite, then the single-GPU measurement set again
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_run56_pytest.log; cat gpurun_out/r2_run56_pytest.log
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py 2>&1 | tail -1
bash tools/r2_run39.sh
