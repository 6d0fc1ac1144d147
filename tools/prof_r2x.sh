#!/bin/bash
# round-2 session X: schedule probe of the sub-batch pipeline (head share, explicit cuts, correlation grid cap)
# with the device timeline of one call per setting (B2_PIPE_TRACE)
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
PIPE_PROBE=schedule timeout 400 python tools/pipeline_probe.py 256 > gpurun_out/r2x_pipeline_schedule.txt 2> gpurun_out/r2x_pipeline_trace.txt
cat gpurun_out/r2x_pipeline_schedule.txt
grep -c "b2 pipe" gpurun_out/r2x_pipeline_trace.txt
