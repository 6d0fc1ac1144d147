#!/bin/bash
# Round-2 multi-GPU session: usage (under gpurun --gpus N): bash tools/prof_r2_multi.sh N
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_${N}gpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo ==== BENCH N=$N
timeout 1500 $TR --master-port 29501 bench.py --gpus $N --steps 5 --warmup 3 2> gpurun_out/r2_bench_${N}gpu_stderr.txt | tail -1 > gpurun_out/r2_bench_${N}gpu.json
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_${N}gpu.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "per_rank", "verified_offsets")})
print(d["e2e"]); print(d["config"]["workload"])
PY
tail -3 gpurun_out/r2_bench_${N}gpu_stderr.txt
echo ==== CANDIDATE MODE
timeout 600 $TR --master-port 29502 tools/candidate_mode_bench.py --pairs 1 --ratios 7 2>/dev/null | tail -1 | tee gpurun_out/r2_candidate_mode_${N}gpu_b1.json
timeout 600 $TR --master-port 29503 tools/candidate_mode_bench.py --pairs 4 --ratios 7 2>/dev/null | tail -1 | tee gpurun_out/r2_candidate_mode_${N}gpu_b4.json
echo ==== SWEEP
if [ "$N" = "8" ]; then CELLS="120:4096,120:8192,10:8192,240:4096,30:4096"; else CELLS="120:512,120:4096,10:4096"; fi
timeout 1800 $TR --master-port 29504 tools/sweep.py --cells $CELLS > gpurun_out/r2_sweep_${N}gpu.md 2> gpurun_out/r2_sweep_${N}gpu_err.txt
cat gpurun_out/r2_sweep_${N}gpu.md; tail -3 gpurun_out/r2_sweep_${N}gpu_err.txt
