#!/bin/bash
# round-2 GPU call 3 (1 GPU): test files in separate processes, both bench arms, ncu --set full of the per-Gaussian kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
: > $O/r02c_pytest_gpu.log
for f in tests/test_compose_gpu.py tests/test_losses_gpu.py tests/test_parity_gpu.py; do
  echo "=== $f" >> $O/r02c_pytest_gpu.log
  timeout 900 python -m pytest $f -m gpu -q >> $O/r02c_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r02c_pytest_gpu.log
done
grep -E "^===|passed|failed|^FAILED|rc=" $O/r02c_pytest_gpu.log | head -40
timeout 300 python bench.py > $O/r02c_bench_sgr.json 2> $O/r02c_bench_sgr.err
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/r02c_bench_sgr.json').read().strip().split('\n')[-1])
    print({k:j.get(k) for k in ('value','ms_per_step','gpu_launches')}, j.get('e2e'), j.get('e2e_full_upload',{}).get('ms_per_step'), j['config'].get('stage_ms'))
except Exception as e: print('bench parse failed', e); print(open('gpurun_out/r02c_bench_sgr.err').read()[-1500:])
PY
timeout 300 python bench.py --impl reference > $O/r02c_bench_ref.json 2> $O/r02c_bench_ref.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"preprocess_fwd_kernel|preprocess_bwd_kernel|compose_fwd_kernel|compose_bwd_kernel" -s 8 -c 8 -o $O/r02c_prof_pergauss python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02c_ncu.log 2>&1
SGR_NO_TMA=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"preprocess_fwd_kernel" -s 4 -c 1 -o $O/r02c_prof_fwd_notma python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02c_ncu2.log 2>&1
ls -la $O/*.ncu-rep
echo done
