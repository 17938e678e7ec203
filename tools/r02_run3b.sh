#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
: > $O/r02d_pytest_gpu.log
for f in tests/test_losses_gpu.py tests/test_parity_gpu.py; do
  echo "=== $f" >> $O/r02d_pytest_gpu.log
  timeout 900 python -m pytest $f -m gpu -q >> $O/r02d_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r02d_pytest_gpu.log
done
grep -E "^===|passed|failed|^FAILED|rc=|AssertionError: " $O/r02d_pytest_gpu.log | head -40
echo done
