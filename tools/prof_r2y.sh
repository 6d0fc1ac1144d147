#!/bin/bash
# round-2 session Y: B2_DEVICE_RESIDENT (calls chained across the batch boundary): GPU suite incl. the chained-call
# parity test, smoke, probe (resident vs ordered, a few settings), bench default (resident) and --ordered-calls
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -6 | tee gpurun_out/r2y_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2y_smoke.txt
PIPE_PROBE=resident timeout 300 python tools/pipeline_probe.py 256 2>&1 | tail -12 | tee gpurun_out/r2y_pipeline_resident.txt
timeout 300 python bench.py 2>&1 | tail -1 | tee gpurun_out/r2y_bench.json | cut -c1-300
timeout 300 python bench.py --ordered-calls --no-cpu-baseline --no-oracle-check 2>&1 | tail -1 | tee gpurun_out/r2y_bench_ordered_calls.json | cut -c1-300
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-oracle-check 2>&1 | tail -1 | tee gpurun_out/r2y_bench_20steps.json | cut -c1-300
