#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LDB200_LIB=$PWD/lightdock-rust_b200/variants/lib_rg_tma.so timeout 900 python -m pytest tests/test_gpu_rigid_path.py -m gpu -x -q -k "flex or pose" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_rigid_path.py -m gpu -x -q -k "flex" 2>&1 | tail -3
for lib in "" lightdock-rust_b200/variants/lib_rg_*.so; do
  if [ -n "$lib" ]; then export LDB200_LIB=$PWD/$lib; else unset LDB200_LIB; fi
  timeout 600 python tools/ab_rigid.py 2>&1 | tail -5
done > gpurun_out/r2_run31_ab.log 2>&1
cat gpurun_out/r2_run31_ab.log
