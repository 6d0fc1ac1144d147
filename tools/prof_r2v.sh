#!/bin/bash
# round-2 session V: pipeline on by default: GPU suite, smoke, bench (with oracle check), small-batch break-even
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -6 | tee gpurun_out/r2v_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2v_smoke.txt
timeout 400 python bench.py 2>&1 | tail -1 | tee gpurun_out/r2v_bench.json | cut -c1-400
for b in 96 128; do timeout 120 python tools/pipeline_probe.py $b 2>&1 | tail -6 | tee gpurun_out/r2v_pipeline_${b}pairs.txt; done
