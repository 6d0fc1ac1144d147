#!/bin/bash
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -k "wide or unmasked or four_hours or shifted or adversarial or small_cases or maxscore" 2>&1 | tail -4 | tee gpurun_out/r2k_pytest_big.txt
for cfg in "32 120" "16 240" "128 10"; do
  set -- $cfg
  BIG_BENCH_WS= timeout 600 python tools/big_path_bench.py $1 $2 2>&1 | tail -1 | tee gpurun_out/r2k_big_path_$2min.txt | cut -c1-500
done
