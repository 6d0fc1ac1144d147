#!/bin/bash
# round-2 session W (final tree): GPU suite, smoke, bench; ncu evidence for the lane-per-window VAD kernel
# (full GPU and the 80-SM partitioned shape), launch list of the pipelined 256-pair step; sanitizer over
# the lane kernel / pipeline shapes; reference arm; 512 pairs (per-GPU share of configs[3]) on one GPU
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -6 | tee gpurun_out/r2w_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2w_smoke.txt
timeout 300 python bench.py 2>&1 | tail -1 | tee gpurun_out/r2w_bench.json | cut -c1-300
echo ==== NCU
timeout 200 ncu --set full --clock-control none --import-source on -k regex:vad_lane -s 2 -c 1 -o gpurun_out/r2w_vad_lane -f \
  python tools/vad_partition_ncu.py 16 > gpurun_out/r2w_vad_lane.log 2>&1
B2_VAD_GRID=80 timeout 200 ncu --set full --clock-control none -k regex:vad_lane -s 2 -c 1 -o gpurun_out/r2w_vad_lane_x80 -f \
  python tools/vad_partition_ncu.py 16 > gpurun_out/r2w_vad_lane_x80.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2w_launches_256pairs.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-oracle-check > gpurun_out/r2w_bench_under_ncu.log 2>&1
echo ==== SANITIZER
timeout 240 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/r2w_sanitizer_memcheck.txt 2>&1; tail -3 gpurun_out/r2w_sanitizer_memcheck.txt
timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/r2w_sanitizer_racecheck.txt 2>&1; tail -3 gpurun_out/r2w_sanitizer_racecheck.txt
echo ==== REFERENCE ARM, 512 pairs
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | tee gpurun_out/r2w_bench_reference_arm.json | cut -c1-300
timeout 300 python bench.py --pairs 512 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/r2w_bench_512pairs.json | cut -c1-300
ls -la gpurun_out
