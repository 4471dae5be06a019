#!/bin/bash
# round-2 GPU run 3: full GPU suite on the FLEX / device-cell-list / sharding work, smoke, default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_run3_pytest.log
tail -8 gpurun_out/r2_run3_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_run3_smoke.log 2>&1; tail -5 gpurun_out/r2_run3_smoke.log
timeout 900 python bench.py > gpurun_out/r2_run3_bench.json 2> gpurun_out/r2_run3_bench.err; tail -c 3000 gpurun_out/r2_run3_bench.json; tail -5 gpurun_out/r2_run3_bench.err
