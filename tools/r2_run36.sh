#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rigid_path.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for lib in "" lightdock-rust_b200/variants/lib_*.so; do
  if [ -n "$lib" ]; then export LDB200_LIB=$PWD/$lib; else unset LDB200_LIB; fi
  timeout 600 python tools/ab_rigid.py 2>&1 | tail -11
done > gpurun_out/r2_run36_ab.log 2>&1
cat gpurun_out/r2_run36_ab.log
