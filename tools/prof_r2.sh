#!/bin/bash
# Round-2 GPU session A: parity tests, default bench line, SM-partition probe, sanitizer logs,
# VAD traffic capture, CPU worker sweep.   usage (under gpurun): bash tools/prof_r2.sh
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc; cat /sys/fs/cgroup/cpu.max; python -c "import os; print(len(os.sched_getaffinity(0)))"
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r2a_pytest_gpu.txt
echo ==== BENCH
timeout 1200 python bench.py --steps 5 --warmup 3 2> gpurun_out/r2a_bench_stderr.txt | tail -1 | tee gpurun_out/r2a_bench_default_1gpu.json
tail -5 gpurun_out/r2a_bench_stderr.txt
echo ==== PARTITION PROBE
timeout 900 python tools/partition_probe.py 74 2>&1 | tee gpurun_out/r2a_partition_probe.txt
echo ==== SANITIZER
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_smoke.py > gpurun_out/r2a_sanitizer_$tool.txt 2>&1
  tail -4 gpurun_out/r2a_sanitizer_$tool.txt
done
echo ==== NCU VAD TRAFFIC
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vad_energy -s 4 -c 1 -o gpurun_out/r2_vad -f \
  python bench.py --pairs 16 --steps 1 --warmup 1 --no-cpu-baseline --no-oracle-check > gpurun_out/r2_vad_ncu.log 2>&1
tail -2 gpurun_out/r2_vad_ncu.log
echo ==== CPU SWEEP
timeout 900 python tools/cpu_sweep.py 8 16 32 64 128 > gpurun_out/r2a_cpu_sweep.json
cat gpurun_out/r2a_cpu_sweep.json | head -40
ls -la gpurun_out
