#!/bin/bash
# Builds A/B variants of the CUDA library into lightdock-rust_b200/variants/ (git-ignored *.so; they travel to the GPU
# box with the snapshot).  Usage: tools/build_variants.sh name1:"-DFLAG=.." name2:"..."; load one with LDB200_LIB=<path>.
set -e
cd "$(dirname "$0")/../lightdock-rust_b200"
mkdir -p variants
if [ "$1" = "--clean" ]; then rm -f variants/*; shift; fi
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 \
       -Xcompiler -fPIC,-fvisibility=default,-pthread -I../include $flags -Xptxas -v -shared \
       -o variants/lib_$name.so csrc/ld_capi.cu csrc/ld_probe.cu 2> variants/ptxas_$name.log &
done
wait
grep -c "spill" variants/ptxas_*.log >/dev/null; ls -la variants/*.so
