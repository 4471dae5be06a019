#!/bin/bash
# round-2 GPU run 50: the default bench line (GSO legs at the 100 steps of BASELINE configs[4])
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_run57_bench.json 2> gpurun_out/r2_run57_bench.err; tail -c 400 gpurun_out/r2_run57_bench.json; tail -3 gpurun_out/r2_run57_bench.err
