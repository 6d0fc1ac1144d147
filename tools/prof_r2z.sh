#!/bin/bash
# round-2 session Z: sustained (3 s) runs of the chained and the ordered call mode under the same power regime,
# 512 pairs (per-GPU share of configs[3]) chained, ncu launch list of the chained bench step
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 200 python bench.py --min-seconds 3 --no-cpu-baseline --no-oracle-check 2>/dev/null | tail -1 | tee gpurun_out/r2z_bench_sustained_3s.json | cut -c1-250
timeout 200 python bench.py --min-seconds 3 --ordered-calls --no-cpu-baseline --no-oracle-check 2>/dev/null | tail -1 | tee gpurun_out/r2z_bench_sustained_3s_ordered.json | cut -c1-250
timeout 200 python bench.py --pairs 512 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/r2z_bench_512pairs.json | cut -c1-250
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2z_launches_256pairs.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-oracle-check > gpurun_out/r2z_bench_under_ncu.log 2>&1
