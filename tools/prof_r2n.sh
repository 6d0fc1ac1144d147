#!/bin/bash
# round-2 session N: validation of the final tree (full GPU suite, smoke, bench, big-path bench)
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -4 | tee gpurun_out/r2n_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2n_smoke.txt
timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/r2n_bench.json | cut -c1-600
for cfg in "32 120" "16 240" "128 10"; do
  set -- $cfg
  timeout 600 python tools/big_path_bench.py $1 $2 2>&1 | tail -1 | tee gpurun_out/r2n_big_path_$2min.txt | cut -c1-500
done
