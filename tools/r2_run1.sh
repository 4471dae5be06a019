#!/bin/bash
# round-2 GPU run 1: full GPU test-suite on the new DNA kernel, then A/B timings of the kernel variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_run1_pytest.log
tail -5 gpurun_out/r2_run1_pytest.log
for v in "" dna_c4_b1 dna_c3_b1 dna_c2_b1 dna_c3_b0 dna_c2_b0 dna_t128_c6; do
  if [ -n "$v" ]; then export LDB200_LIB=$PWD/lightdock-rust_b200/variants/lib_$v.so; else unset LDB200_LIB; fi
  echo "== DNA variant ${v:-default}"
  timeout 300 python tools/dna_timing.py 2>&1 | tail -4
done > gpurun_out/r2_run1_dna.log 2>&1
cat gpurun_out/r2_run1_dna.log
for v in "" rg768 rg704; do
  if [ -n "$v" ]; then export LDB200_LIB=$PWD/lightdock-rust_b200/variants/lib_$v.so; else unset LDB200_LIB; fi
  echo "== rigid variant ${v:-default}"
  SWARMS=100 timeout 600 python tools/rigid_bringup.py 2>&1 | grep -E "^path 2|rigid vs"
done > gpurun_out/r2_run1_rigid.log 2>&1
cat gpurun_out/r2_run1_rigid.log
unset LDB200_LIB
timeout 600 python tools/config_timings.py > gpurun_out/r2_run1_configs.log 2>&1; cat gpurun_out/r2_run1_configs.log
