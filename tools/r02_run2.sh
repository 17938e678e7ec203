#!/bin/bash
# round-2 GPU call 2 (1 GPU): sanitizer on the composer edge test, full GPU test suite, bench A/B of the new kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_compose_gpu.py -x -q -m gpu -k edge > $O/r02_sanitize_compose.log 2>&1
grep -E "Invalid|ERROR SUMMARY|at .*\+0x|by thread|Address" $O/r02_sanitize_compose.log | head -20
timeout 900 python -m pytest tests -m gpu -q > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu.log
tail -15 $O/r02_pytest_gpu.log
timeout 300 python bench.py > $O/r02b_bench_sgr.json 2> $O/r02b_bench_sgr.err
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/r02b_bench_sgr.json').read().strip().split('\n')[-1])
    print({k:j.get(k) for k in ('value','ms_per_step','gpu_launches')}, j.get('e2e'), j['config'].get('stage_ms'))
except Exception as e: print('bench parse failed', e); print(open('gpurun_out/r02b_bench_sgr.err').read()[-2000:])
PY
for v in "SGR_NO_TMA=1" "SGR_BWD2_BATCH=8" "SGR_BWD2_BATCH=32"; do
  echo "== $v"; env $v timeout 200 python bench.py --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['config']['stage_ms'])
except Exception as e: print('fail', e)"
done
timeout 300 python bench.py --impl reference > $O/r02b_bench_ref.json 2> $O/r02b_bench_ref.err
tail -c 900 $O/r02b_bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02b_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/r02b_launch_bench.log 2>&1
echo done
