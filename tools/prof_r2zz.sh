#!/bin/bash
# round-2 final-tree validation: GPU suite, smoke, default bench line
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -4 | tee gpurun_out/r2zz_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2zz_smoke.txt
timeout 200 python bench.py 2>/dev/null | tail -1 | tee gpurun_out/r2zz_bench.json | cut -c1-200
