#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
: > $O/r02f_pytest_gpu.log
for f in tests/test_compose_gpu.py tests/test_losses_gpu.py tests/test_parity_gpu.py; do
  echo "=== $f" >> $O/r02f_pytest_gpu.log
  timeout 900 python -m pytest $f -m gpu -q >> $O/r02f_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r02f_pytest_gpu.log
done
grep -E "^===|passed|failed|^FAILED|rc=|AssertionError: " $O/r02f_pytest_gpu.log | cut -c1-300 | head -40
timeout 300 python bench.py > $O/r02f_bench_sgr.json 2> $O/r02f_bench_sgr.err
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/r02f_bench_sgr.json').read().strip().split('\n')[-1])
    print({k:j.get(k) for k in ('value','ms_per_step','gpu_launches')}, j.get('e2e'), j.get('e2e_full_upload',{}).get('ms_per_step'), j['config'].get('stage_ms'))
except Exception as e: print('bench parse failed', e); print(open('gpurun_out/r02f_bench_sgr.err').read()[-1500:])
PY
echo done
