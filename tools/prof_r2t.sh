#!/bin/bash
# round-2 session T: lane kernel with two pipelines per CTA: ladder (correctness), batch sweep
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 150 python tools/lane_debug.py 2>&1 | tail -14 | tee gpurun_out/r2t_ladder.txt
timeout 200 python tools/lane_probe.py 74 2>&1 | tee gpurun_out/r2t_lane_probe.txt | tail -40
