#!/bin/bash
# bench.py at N GPUs the way the driver launches it; usage: tools/r2_scale.sh N
cd "$(dirname "$0")/.."
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -c 2500 gpurun_out/r2_bench_${N}gpu.json; tail -3 gpurun_out/r2_bench_${N}gpu.err
