#!/bin/bash
# Round-2 GPU session G (1 GPU): big-path timing after the fixes; configs[3]'s per-GPU footprint (512 pairs) on one GPU.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
for cfg in "32 120" "16 240" "64 30" "128 10"; do
  set -- $cfg
  timeout 600 python tools/big_path_bench.py $1 $2 2>&1 | tail -1 | tee gpurun_out/r2g_big_path_$2min.txt | cut -c1-700
done
timeout 1500 python bench.py --pairs 512 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2g_bench512_stderr.txt | tail -1 > gpurun_out/r2g_bench_512pairs_1gpu.json
cut -c1-300 gpurun_out/r2g_bench_512pairs_1gpu.json; tail -3 gpurun_out/r2g_bench512_stderr.txt
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
timeout 900 python -m pytest tests -m gpu -q --tb=short -k "adversarial or auditok or wide" 2>&1 | tail -5
