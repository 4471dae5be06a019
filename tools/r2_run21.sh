#!/bin/bash
# round-2 GPU run 21: poses per CTA of transform_kernel (DNA 1azp: transform is 13 % of the pair time)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in "" lightdock-rust_b200/variants/lib_tr_*.so; do
  if [ -n "$lib" ]; then export LDB200_LIB=$PWD/$lib; else unset LDB200_LIB; fi
  echo "== ${lib:-default (4 poses per CTA)}"
  timeout 300 python tools/dna_timing.py 2>&1 | grep -E "parity|poses " | tail -3
done > gpurun_out/r2_run21_transform.log 2>&1
cat gpurun_out/r2_run21_transform.log
