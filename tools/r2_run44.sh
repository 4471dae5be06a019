#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in "" lightdock-rust_b200/variants/lib_flex*.so; do
  if [ -n "$lib" ]; then export LDB200_LIB=$PWD/$lib; else unset LDB200_LIB; fi
  timeout 600 python tools/ab_rigid.py 2>&1 | grep -E "1ppe|2uuy|1czy|80000"
done > gpurun_out/r2_run44_flexwarps.log 2>&1
cat gpurun_out/r2_run44_flexwarps.log
