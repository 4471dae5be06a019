#!/bin/bash
# round-2 GPU run 17: randomized parity campaign on the shipped kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
POSES=400000 FLEX_POSES=100000 timeout 2400 python tools/parity_campaign.py > gpurun_out/r2_parity_campaign.log 2>&1
tail -12 gpurun_out/r2_parity_campaign.log
