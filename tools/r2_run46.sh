#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trajectory.py -m gpu -x -q -k "dna or azp or pydock or thresholds or splits or determin or edge" 2>&1 | tail -3
timeout 300 python tools/dna_timing.py 2>&1 | grep -E "parity|poses |fractions" | tail -5 > gpurun_out/r2_run46_dna.log 2>&1
cat gpurun_out/r2_run46_dna.log
