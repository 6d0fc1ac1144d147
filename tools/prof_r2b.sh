#!/bin/bash
# Round-2 GPU session B (1 GPU): full parity suite, sanitizer round 2, ncu evidence, sustained run, sweep.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r2b_pytest_gpu.txt
echo ==== SANITIZER
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/r2b_sanitizer_racecheck.txt 2>&1; tail -3 gpurun_out/r2b_sanitizer_racecheck.txt
B2_ACC=reg timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_smoke.py > gpurun_out/r2b_sanitizer_synccheck_regacc.txt 2>&1; tail -3 gpurun_out/r2b_sanitizer_synccheck_regacc.txt
B2_ACC=reg timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/r2b_sanitizer_memcheck_regacc.txt 2>&1; tail -3 gpurun_out/r2b_sanitizer_memcheck_regacc.txt
timeout 900 compute-sanitizer --tool initcheck python tools/sanitize_smoke.py > gpurun_out/r2b_sanitizer_initcheck.txt 2>&1; tail -3 gpurun_out/r2b_sanitizer_initcheck.txt
echo ==== NCU
B2_VAD_CONSUMERS=512 B2_VAD_CTAS_FORCE=1 B2_VAD_STAGES=5 B2_VAD_GRID=74 timeout 600 ncu --set full --clock-control none -k regex:vad_energy -s 2 -c 1 -o gpurun_out/r2_vad_x74 -f python tools/vad_partition_ncu.py 16 > gpurun_out/r2_vad_x74.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sub_correlate_bits|ref_spectra' -s 4 -c 2 -o gpurun_out/r2_corr_fullload -f \
  python bench.py --pairs 148 --steps 1 --warmup 1 --no-cpu-baseline --no-oracle-check > gpurun_out/r2_corr_fullload.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2b_launches_256pairs.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-oracle-check > gpurun_out/r2b_bench_under_ncu.log 2>&1
echo ==== SUSTAINED
timeout 900 python bench.py --steps 5 --warmup 3 --min-seconds 3 --no-cpu-baseline --no-oracle-check 2>/dev/null | tail -1 | tee gpurun_out/r2b_bench_sustained_3s.json | head -c 600
echo
echo ==== BENCH reference arm
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 | tail -1 | tee gpurun_out/r2b_bench_reference_arm.json | head -c 400
echo
echo ==== SWEEP
timeout 2400 python tools/sweep.py > gpurun_out/r2b_sweep_1gpu.md 2> gpurun_out/r2b_sweep_err.txt
cat gpurun_out/r2b_sweep_1gpu.md; tail -3 gpurun_out/r2b_sweep_err.txt
ls -la gpurun_out
