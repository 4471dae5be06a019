#!/bin/bash
# round-2 GPU run 18: DNA/pyDock kernel, warps per CTA
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in 0 8 7 6 5 4 3; do echo "== LDB200_DNA_WARPS=$w"; LDB200_DNA_WARPS=$w timeout 300 python tools/dna_timing.py 2>&1 | grep -E "parity|poses" | tail -3; done > gpurun_out/r2_run18_dna.log 2>&1
cat gpurun_out/r2_run18_dna.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "dna or azp or pydock" 2>&1 | tail -3
