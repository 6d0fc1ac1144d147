#!/bin/bash
# Round-2 GPU session I (1 GPU): what the driver runs at round end, on the final tree, plus the pageable-input latency.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); print('build ok')" 2>&1 | tail -1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2i_pytest_gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 | tail -1 > gpurun_out/r2i_bench_reference_arm.json; cut -c1-200 gpurun_out/r2i_bench_reference_arm.json
timeout 1200 python bench.py --gpus 1 --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2i_bench_default_1gpu.json; cut -c1-250 gpurun_out/r2i_bench_default_1gpu.json
timeout 600 python tools/latency_probe.py 2>&1 | tail -1 | tee gpurun_out/r2i_latency_single_pair.json | cut -c1-700
timeout 600 compute-sanitizer --tool memcheck python tools/latency_probe.py > gpurun_out/r2i_sanitizer_memcheck_latency.txt 2>&1; tail -2 gpurun_out/r2i_sanitizer_memcheck_latency.txt
