#!/bin/bash
# round-2 GPU run 10: FLEX staging by per-warp bulk copy: tests + A/B against the previous build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out


for lib in "" lightdock-rust_b200/variants/lib_rg_*.so; do
  if [ -n "$lib" ]; then export LDB200_LIB=$PWD/$lib; else unset LDB200_LIB; fi
  timeout 600 python tools/ab_rigid.py 2>&1 | tail -6
done > gpurun_out/r2_run27_ab.log 2>&1
cat gpurun_out/r2_run27_ab.log
