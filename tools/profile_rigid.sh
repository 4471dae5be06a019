#!/bin/bash
# ncu capture of the rigid DFIRE kernel on a 60-swarm slice of the bench workload (run under gpurun).
set -e
TAG=${1:-rigid}
export SWARMS=${SWARMS:-60}
python tools/rigid_bringup.py > gpurun_out/bringup_$TAG.log 2>&1 || true
tail -8 gpurun_out/bringup_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:dfire_rigid_kernel -s 9 -c 1 \
    -o gpurun_out/prof_$TAG -f python tools/rigid_bringup.py > gpurun_out/ncu_$TAG.log 2>&1 || true
tail -3 gpurun_out/ncu_$TAG.log
