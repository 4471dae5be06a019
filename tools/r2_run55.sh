#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in "" lightdock-rust_b200/variants/lib_dna_*.so; do
  if [ -n "$lib" ]; then export LDB200_LIB=$PWD/$lib; else unset LDB200_LIB; fi
  echo "== ${lib:-default}"
  timeout 300 python tools/dna_timing.py 2>&1 | grep -E "parity|poses " | tail -3
done > gpurun_out/r2_run55_dna.log 2>&1
cat gpurun_out/r2_run55_dna.log
