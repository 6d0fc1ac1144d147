#!/bin/bash
# Round-2 GPU session H (1 GPU): final validation of the tree.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15 | tee gpurun_out/r2h_pytest_gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for cfg in "32 120" "16 240" "128 10"; do
  set -- $cfg
  timeout 600 python tools/big_path_bench.py $1 $2 2>&1 | tail -1 | tee gpurun_out/r2h_big_path_$2min.txt | cut -c1-500
done
BIG_BENCH_PATHS=big BIG_BENCH_WS= timeout 300 python tools/big_path_bench.py 1 120 2>&1 | tail -1 | tee gpurun_out/r2h_big_path_1pair.txt | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2h_launches_bigpath_1pair.csv \
  env BIG_BENCH_PATHS=big BIG_BENCH_WS= python tools/big_path_bench.py 1 120 > /dev/null 2>&1
timeout 1200 python bench.py --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2h_bench_default_1gpu.json; cut -c1-250 gpurun_out/r2h_bench_default_1gpu.json
