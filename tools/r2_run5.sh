#!/bin/bash
# round-2 GPU run 5: A/B of the rigid-kernel variants in lightdock-rust_b200/variants/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in "" lightdock-rust_b200/variants/lib_rg_*.so; do
  if [ -n "$lib" ]; then export LDB200_LIB=$PWD/$lib; else unset LDB200_LIB; fi
  timeout 600 python tools/ab_rigid.py 2>&1 | tail -9
done > gpurun_out/r2_run6_ab.log 2>&1
cat gpurun_out/r2_run6_ab.log
