#!/bin/bash
# round-2 session Q: size ladder of the lane kernel against the lane-group kernel and a torch yardstick
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 150 python tools/lane_debug.py 2>&1 | tail -20 | tee gpurun_out/r2q_ladder.txt
timeout 200 python tools/lane_probe.py 74 2>&1 | tee gpurun_out/r2q_lane_probe.txt | tail -40
