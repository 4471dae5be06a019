#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(for lib in "" lightdock-rust_b200/variants/lib_stcs.so; do
  if [ -n "$lib" ]; then export LDB200_LIB=$PWD/$lib; else unset LDB200_LIB; fi
  for c in 0 0.75; do echo "== lib=${lib:-default} LDB200_CELL=$c"; LDB200_CELL=$c timeout 600 python tools/ab_rigid.py 2>&1 | grep -E "1k4c (10000|80000)|rank 0|1ppe|2uuy"; done
done) > gpurun_out/r2_run48_stcs.log 2>&1
cat gpurun_out/r2_run48_stcs.log
