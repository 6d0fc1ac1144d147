#!/bin/bash
# Round-2 GPU session L: the GPU parity suite under compute-sanitizer memcheck (the tests that fit its 10-100x slowdown).
set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r2l_pytest_gpu.txt
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x \
  -k "not config3 and not 65535 and not two_hour and not four_hours and not shifted and not adversarial and not wide_window and not unmasked_batch and not one_long_call and not sync_two_hour and not gss_batched" \
  > gpurun_out/r2l_sanitizer_memcheck_pytest.txt 2>&1
echo "memcheck pytest exit code: $?" | tee -a gpurun_out/r2l_sanitizer_memcheck_pytest.txt
tail -6 gpurun_out/r2l_sanitizer_memcheck_pytest.txt
