#!/bin/bash
# round-2 1-GPU call: A/B of blend_bwd2 batch-size / occupancy variants (parity + time), e2e leg as a CUDA graph
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
for v in 32 16 165 323; do
  echo "== SGR_BWD2_BATCH=$v"
  SGR_BWD2_BATCH=$v timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "golden or live_reference or oracle_small or config_C" 2>&1 | tail -2 | cut -c1-200
  SGR_BWD2_BATCH=$v timeout 200 python bench.py --no-e2e --no-cpu-baseline 2>$O/r02j_bench_$v.err | tail -1 > $O/r02j_bench_$v.json
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/r02j_bench_$v.json').read()); print('ms/step', round(j['ms_per_step'],4), j['config']['stage_ms'])
except Exception as e: print('fail', e); print(open('gpurun_out/r02j_bench_$v.err').read()[-600:])
PY
done
timeout 400 python bench.py --no-cpu-baseline > $O/r02j_bench_sgr.json 2> $O/r02j_bench_sgr.err
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/r02j_bench_sgr.json').read().strip().split('\n')[-1])
    print({k:j.get(k) for k in ('value','ms_per_step','gpu_launches')}, j.get('e2e'), j.get('e2e_full_upload',{}).get('ms_per_step'))
except Exception as e: print('bench parse failed', e); print(open('gpurun_out/r02j_bench_sgr.err').read()[-1500:])
PY
echo done
