#!/bin/bash
# round-2 session R: integer pipe microbenchmark (which byte/halfword dot products are fast on B200)
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 60 tools/microbench/int_pipes 2>&1 | tee gpurun_out/r2r_int_pipes.txt
