#!/bin/bash
# round-2 GPU run 45: GPU test-suite and the default bench line of the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_run45_pytest.log; cat gpurun_out/r2_run45_pytest.log
timeout 900 python bench.py > gpurun_out/r2_run45_bench.json 2> gpurun_out/r2_run45_bench.err; tail -c 600 gpurun_out/r2_run45_bench.json; tail -3 gpurun_out/r2_run45_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_run45_ref.json 2> gpurun_out/r2_run45_ref.err; tail -c 700 gpurun_out/r2_run45_ref.json
