#!/bin/bash
# Round-2 GPU session E (1 GPU): everything that still needs a GPU, most important first.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r2e_pytest_gpu.txt
echo ==== BIG PATH
for cfg in "32 120" "16 240" "64 30" "128 10"; do
  set -- $cfg
  timeout 600 python tools/big_path_bench.py $1 $2 > gpurun_out/r2e_big_path_$2min.txt 2>&1
  tail -1 gpurun_out/r2e_big_path_$2min.txt | cut -c1-900
done
echo ==== BENCH
timeout 1200 python bench.py --steps 5 --warmup 3 2> gpurun_out/r2e_bench_stderr.txt | tail -1 > gpurun_out/r2e_bench_default_1gpu.json
cut -c1-400 gpurun_out/r2e_bench_default_1gpu.json; tail -3 gpurun_out/r2e_bench_stderr.txt
echo ==== LATENCY
timeout 600 python tools/latency_probe.py 2>&1 | tail -1 | tee gpurun_out/r2e_latency_single_pair.json
echo ==== NCU corr
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sub_correlate_bits -s 1 -c 1 -o gpurun_out/r2_subcorr_fullload -f \
  python bench.py --pairs 148 --steps 1 --warmup 1 --no-cpu-baseline --no-oracle-check > gpurun_out/r2_subcorr_fullload.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'big_cols|big_rows' -s 10 -c 5 -o gpurun_out/r2_bigfft -f \
  python tools/big_path_bench.py 8 120 > gpurun_out/r2_bigfft_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2e_launches_bigpath.csv \
  env BIG_BENCH_PATHS=big BIG_BENCH_WS= python tools/big_path_bench.py 8 120 > /dev/null 2>&1
echo ==== SANITIZER big path
B2_ACC=reg BIG_BENCH_PATHS=big BIG_BENCH_WS=256 timeout 900 compute-sanitizer --tool memcheck python tools/big_path_bench.py 3 30 > gpurun_out/r2e_sanitizer_memcheck_bigpath.txt 2>&1
tail -3 gpurun_out/r2e_sanitizer_memcheck_bigpath.txt
ls -la gpurun_out | tail -20
