#!/bin/bash
# round-2 GPU run 14: whole GPU suite after the conformations tool / gso slab / CLI exit changes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_run14_pytest.log
tail -15 gpurun_out/r2_run14_pytest.log
