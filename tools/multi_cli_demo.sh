#!/bin/bash
# Times lightdock-rust-multi on N synthetic 1k4c swarms x 100 steps (run under gpurun; GPUs = all visible).
N=${1:-64}
ROOT=${GRAFT_REPO_ROOT:-$(pwd)}
W=/tmp/mrun_$N; rm -rf $W; mkdir -p $W/init; cd $W
export LIGHTDOCK_DATA=$(python - <<PY
import sys
sys.path.insert(0, "$ROOT/lightdock-rust_b200")
from ldb200 import workload
import numpy as np
dc, _ = workload.ensure_dcparams_dir("$W")
p = workload.synthetic_1k4c_swarms($N, 200)
for s in range($N):
    np.savetxt(f"init/initial_positions_{s}.dat", p[s], fmt="%.17g")
print(dc)
PY
)
time $ROOT/lightdock-rust_b200/bin/lightdock-rust-multi $ROOT/tests/golden/1k4c/setup.json 100 dfire init/initial_positions_*.dat 2>&1 | tail -3
ls swarm_5 | tr '\n' ' '; echo
time $ROOT/lightdock-rust_b200/bin/lightdock-rust $ROOT/tests/golden/1k4c/setup.json init/initial_positions_5.dat 100 dfire > /dev/null
