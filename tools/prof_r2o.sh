#!/bin/bash
# round-2 session O: lane-per-window VAD kernel - parity, throughput vs SM count, partition overlap
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "vad or auditok or stream or sync_batch or config3" 2>&1 | tail -6 | tee gpurun_out/r2o_pytest_vad.txt
timeout 600 python tools/lane_probe.py 74 2>&1 | tee gpurun_out/r2o_lane_probe.txt | tail -60
timeout 300 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r2o_bench.json | cut -c1-300
