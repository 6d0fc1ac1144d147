#!/bin/bash
# Round-1 GPU session: parity tests, launch list, full captures of the main kernels, bench.
# usage (from the repo root, under gpurun):  bash tools/prof_r1.sh [tag] [pairs]
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
TAG=${1:-r1}
PAIRS=${2:-148}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
# launch list (cold-cache, serialised: compare SHARES, not absolutes); skip the warm-up launches
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py --pairs ${PAIRS} --steps 1 --warmup 3 --no-cpu-baseline \
  > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
# full captures of the main kernels (one launch each, after warm-up)
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'sub_correlate|ref_spectra|vad_energy|rescore|raster' -s 6 -c 8 -o gpurun_out/prof_${TAG} -f \
  python bench.py --pairs 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_${TAG}.log 2>&1
ls -la gpurun_out
echo ==== BENCH
timeout 900 python bench.py --pairs ${PAIRS} --steps 5 --warmup 3 ${BENCH_EXTRA} 2>&1 | tail -1 | tee gpurun_out/bench${PAIRS}_${TAG}.json
