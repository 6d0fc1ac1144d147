#!/bin/bash
# Round-2 GPU session D: debug the large-FFT path at 2 h / 4 h batch sizes.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
for cfg in "2 120 big" "8 120 big" "32 120 big" "2 120 tiled"; do
  set -- $cfg
  BIG_BENCH_PATHS=$3 BIG_BENCH_WS= CUDA_LAUNCH_BLOCKING=1 timeout 300 python tools/big_path_bench.py $1 $2 > gpurun_out/r2d_big_$1_$2_$3.txt 2>&1
  tail -4 gpurun_out/r2d_big_$1_$2_$3.txt | cut -c1-600
done
BIG_BENCH_PATHS=big BIG_BENCH_WS= B2_ACC=reg timeout 600 compute-sanitizer --tool memcheck python tools/big_path_bench.py 2 120 > gpurun_out/r2d_memcheck_big_2pairs.txt 2>&1
grep -m 20 -E "Invalid|ERROR|error|at " gpurun_out/r2d_memcheck_big_2pairs.txt | head -30
