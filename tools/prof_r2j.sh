#!/bin/bash
# source-level capture of the large-FFT kernels (where do the cycles go inside a tile?)
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'big_rows_kernel|big_cols' -s 12 -c 3 -o gpurun_out/r2j_bigfft_src -f \
  env BIG_BENCH_PATHS=big BIG_BENCH_WS= python tools/big_path_bench.py 8 120 > gpurun_out/r2j_ncu.log 2>&1
tail -2 gpurun_out/r2j_ncu.log
