#!/bin/bash
# round-2 GPU run 39: final single-GPU measurement set of the kernels as shipped (counters, launch list, bench line, --set full captures)
cd "$(dirname "$0")/.."
bash tools/r2_run9.sh
bash tools/r2_final_profiles.sh
