#!/bin/bash
# round-2 GPU run 15: whole GPU suite (no -x), then steady-state ncu --set full captures of the FLEX kernel (2uuy, 1czy)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_run15_pytest.log
tail -15 gpurun_out/r2_run15_pytest.log
for cfg in 2uuy 1czy; do
  ncu --set full --clock-control none --import-source on -k regex:dfire_rigid_kernel --profile-from-start off -c 1 \
      -o gpurun_out/prof_r2_${cfg}_flex -f python tools/count_target.py $cfg > gpurun_out/ncu_r2_${cfg}_flex.log 2>&1
  tail -2 gpurun_out/ncu_r2_${cfg}_flex.log | cut -c1-200
done
