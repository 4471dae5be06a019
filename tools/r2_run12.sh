#!/bin/bash
# round-2 GPU run 12: work units per SM at 10,000 / 80,000 poses; CLI wall clocks after the fast exit
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for u in 8 16 32 64 128; do LDB200_UNITS_PER_SM=$u timeout 300 python tools/units_sweep.py 2>&1 | tail -2; done > gpurun_out/r2_run12_units.log 2>&1
cat gpurun_out/r2_run12_units.log
python - <<'PY' > gpurun_out/r2_run12_cli.log 2>&1
import json, os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "lightdock-rust_b200"))
import bench
from ldb200 import workload
dc_dir, _ = workload.ensure_dcparams_dir("/tmp")
for rep in range(3):
    out = bench.single_swarm_runs(dc_dir)
    print({k: (round(v["wall_s"], 3), round(v["breakdown"]["process_start_and_exit_ms"]), round(v["device_gso"]["wall_s"], 3)) for k, v in out.items()})
PY
cat gpurun_out/r2_run12_cli.log
