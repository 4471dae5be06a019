#!/bin/bash
# round-2 GPU run 9: kernel counters for the rooflines (-> profiles/r2_kernel_counts.json), launch list of a short
# bench run, then the default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/capture_counts.sh
python tools/make_counts.py
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-single-swarm-runs --gso-steps 3 > gpurun_out/r2_launches_bench.log 2>&1
tail -c 400 gpurun_out/r2_launches_bench.log
timeout 900 python bench.py > gpurun_out/r2_run9_bench.json 2> gpurun_out/r2_run9_bench.err; tail -c 1500 gpurun_out/r2_run9_bench.json; tail -5 gpurun_out/r2_run9_bench.err
cp profiles/r2_kernel_counts.json gpurun_out/r2_kernel_counts.json
