#!/bin/bash
# Round-2 GPU session F (1 GPU): re-validate after the last kernel / host changes.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 | tee gpurun_out/r2f_pytest_gpu.txt
for cfg in "32 120" "16 240" "64 30"; do
  set -- $cfg
  BIG_BENCH_WS= timeout 600 python tools/big_path_bench.py $1 $2 2>&1 | tail -1 | tee gpurun_out/r2f_big_path_$2min.txt | cut -c1-700
done
timeout 600 python tools/latency_probe.py 2>&1 | tail -1 | tee gpurun_out/r2f_latency_single_pair.json
timeout 900 ncu --set full --clock-control none -k regex:'big_cols|big_rows' -s 10 -c 5 -o gpurun_out/r2f_bigfft -f \
  env BIG_BENCH_PATHS=big BIG_BENCH_WS= python tools/big_path_bench.py 8 120 > gpurun_out/r2f_bigfft_ncu.log 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
