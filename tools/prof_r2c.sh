#!/bin/bash
# Round-2 GPU session C: large-FFT path parity + timing.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "wide or unmasked or four_hours or shifted or auditok_detector or small_cases or kats or maxscore" 2>&1 | tail -25 | tee gpurun_out/r2c_pytest_big.txt
timeout 600 python tools/big_path_bench.py 32 120 2>&1 | tail -2 | tee gpurun_out/r2c_big_path_2h.json
timeout 600 python tools/big_path_bench.py 16 240 2>&1 | tail -2 | tee gpurun_out/r2c_big_path_4h.json
timeout 600 python tools/big_path_bench.py 64 30 2>&1 | tail -2 | tee gpurun_out/r2c_big_path_30min.json
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2c_pytest_gpu.txt
B2_ACC=reg timeout 600 compute-sanitizer --tool memcheck python -c "
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests/golden')
import numpy as np, cases
from ffsubsync_b200.aligners import FFTAligner
ref, sub = cases.shifted_pair(60000)
print(FFTAligner().fit_transform(ref, sub, get_score=True))
" > gpurun_out/r2c_sanitizer_memcheck_big.txt 2>&1; tail -3 gpurun_out/r2c_sanitizer_memcheck_big.txt
