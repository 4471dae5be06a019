#!/bin/bash
# Round artefacts for the bench configuration (run under gpurun, 1 GPU):
#   bench line, ncu launch list of the same command, one --set full capture of the dominant kernel.
TAG=${1:-r1}
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 600 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --gso-steps 0 --no-single-swarm-runs > gpurun_out/ncu_launches_$TAG.log 2>&1
tail -2 gpurun_out/ncu_launches_$TAG.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:dfire_rigid_kernel -s 3 -c 1 -f -o gpurun_out/prof_bench_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --gso-steps 0 --no-single-swarm-runs > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log | cut -c1-300
