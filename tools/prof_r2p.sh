#!/bin/bash
# round-2 session P: lane-per-window VAD kernel after the producer rewrite (short timeouts: new kernel)
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 300 python tools/lane_probe.py 74 2>&1 | tee gpurun_out/r2p_lane_probe.txt | tail -40
timeout 240 python -m pytest tests -m gpu -q --tb=short -x -k "vad or auditok or stream" 2>&1 | tail -6 | tee gpurun_out/r2p_pytest_vad.txt
timeout 200 python bench.py --steps 5 --warmup 3 --no-oracle-check 2>&1 | tail -1 | tee gpurun_out/r2p_bench.json | cut -c1-300
