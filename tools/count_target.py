"""One large batch of one configuration through the C ABI, for `ncu -k regex:<kernel> -s <skip> -c 1` captures of the
per-kernel counters bench.py's rooflines use (tools/capture_counts.sh).  Prints the batch size.

    python tools/count_target.py <1czy|1ppe|2uuy|ab_icode|1azp|1k4c_bench> [n_warm]
Two warm-up calls of the same batch (FLEX handles learn their slacks and rebuild their lists in the first), then the
captured call between cudaProfilerStart/Stop (ncu --profile-from-start off -c 1).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("", "oracle", "lightdock-rust_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import oracle as O  # noqa: E402
from helpers import case, scorer_from_oracle  # noqa: E402
from ldb200 import workload  # noqa: E402


def config_poses(name):
    """The 20,000-pose batch of a BASELINE config: its 200 shipped start poses x 100, translations jittered by 1 A
    (seeded), ANM extents kept.  1k4c_bench: 60,000 of the 80,000 poses of the headline workload."""
    if name == "1k4c_bench":
        cx, _, _ = case("1k4c", O.DFIRE)
        # 300 of the 400 swarms (every 4th left out): 60,000 poses, below the size at which ld_score_batch splits a
        # call in two launches, so the captured launch is the whole batch
        sw = workload.synthetic_1k4c_swarms(400, 200)
        keep = np.arange(400) % 4 != 3
        return cx, np.ascontiguousarray(sw[keep].reshape(-1, 7))
    cx, pos, _ = case(name, O.DNA if name == "1azp" else O.DFIRE)
    rng = np.random.default_rng(1)
    big = np.tile(pos, (100, 1))
    big[:, :3] += rng.normal(0, 1.0, size=(len(big), 3))
    return cx, big


if __name__ == "__main__":
    name = sys.argv[1]
    cx, poses = config_poses(name)
    sc = scorer_from_oracle(cx)
    for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
        sc.energy(poses)
    while sc.stats()["pair_launches"] > 1:  # the library splits a call that exceeds its chunk size (FLEX: 64 MB of
        poses = np.ascontiguousarray(poses[: len(poses) // sc.stats()["pair_launches"]])  # ligand blocks) in several
        sc.energy(poses)                                                                    # launches: one launch = the batch
    # only the last call is visible to `ncu --profile-from-start off`: its first launch of the pair kernel is a
    # steady-state launch whatever the number of chunks a call is split in (FLEX: two per 20,000 poses)
    import torch
    rt = torch.cuda.cudart()
    rt.cudaProfilerStart()
    sc.energy(poses)
    rt.cudaProfilerStop()
    print(name, len(poses), "poses", sc.path_info()[:80])
