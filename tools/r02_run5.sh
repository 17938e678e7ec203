#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
: > $O/r02e_pytest_gpu.log
for f in tests/test_compose_gpu.py tests/test_losses_gpu.py tests/test_parity_gpu.py; do
  echo "=== $f" >> $O/r02e_pytest_gpu.log
  timeout 900 python -m pytest $f -m gpu -q >> $O/r02e_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r02e_pytest_gpu.log
done
grep -E "^===|passed|failed|^FAILED|rc=|AssertionError: " $O/r02e_pytest_gpu.log | cut -c1-200 | head -40
for v in "X=1" "SGR_NO_TMA=1"; do
  echo "== $v"; env $v timeout 200 python bench.py --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['config']['stage_ms'])
except Exception as e: print('fail', e)"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02e_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"preprocess_fwd_kernel|preprocess_bwd_tma_kernel" -s 8 -c 2 -o $O/r02e_prof_pergauss python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02e_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"compose_fwd_kernel|compose_bwd_kernel" -c 2 -o $O/r02e_prof_compose python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r02e_ncu2.log 2>&1
ls -la $O/r02e*.ncu-rep
echo done
