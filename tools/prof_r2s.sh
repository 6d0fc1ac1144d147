#!/bin/bash
# round-2 session S: lane kernel v4 (sequence numbers, parallel staging): ladder, probe, VAD tests
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 150 python tools/lane_debug.py 2>&1 | tail -20 | tee gpurun_out/r2s_ladder.txt
timeout 200 python tools/lane_probe.py 74 2>&1 | tee gpurun_out/r2s_lane_probe.txt | tail -40
timeout 240 python -m pytest tests -m gpu -q --tb=short -x -k "vad or auditok or stream" 2>&1 | tail -6 | tee gpurun_out/r2s_pytest_vad.txt
