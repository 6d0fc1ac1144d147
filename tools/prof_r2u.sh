#!/bin/bash
# round-2 session U: sub-batch pipeline with SM partitioning (lane-per-window VAD), 256 pairs
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 300 python tools/pipeline_probe.py 256 2>&1 | tee gpurun_out/r2u_pipeline_probe.txt | tail -30
