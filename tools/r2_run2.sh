#!/bin/bash
# round-2 GPU run: DNA parity tests, A/B timing of the DNA kernel variants in lightdock-rust_b200/variants/, one ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2_dna_v2}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dna or azp" 2>&1 | tail -3
for lib in "" lightdock-rust_b200/variants/lib_dna*.so; do
  if [ -n "$lib" ]; then export LDB200_LIB=$PWD/$lib; else unset LDB200_LIB; fi
  echo "== DNA variant ${lib:-default}"
  timeout 300 python tools/dna_timing.py 2>&1 | tail -3
done > gpurun_out/${TAG}_timing.log 2>&1
cat gpurun_out/${TAG}_timing.log
unset LDB200_LIB
bash tools/r2_ncu_dna.sh $TAG
