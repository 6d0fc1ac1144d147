#!/bin/bash
# Round-2 GPU session M: flakiness check - the GPU parity suite five times in a row on the final tree.
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
for i in 1 2 3 4 5; do
  timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -1 | tee -a gpurun_out/r2m_pytest_gpu_x5.txt
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a gpurun_out/r2m_pytest_gpu_x5.txt
