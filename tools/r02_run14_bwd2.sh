#!/bin/bash
# round-2 1-GPU call: (a) block-run exchange on emulated ranks, (b) blend_bwd2 with chunk-local moments (+ fast-exp A/B): parity, then time
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "fused or capacity or world1 or sharded" > $O/r02k_pytest_fused.log 2>&1; echo "rc=$?" >> $O/r02k_pytest_fused.log
grep -E "passed|failed|^FAILED|rc=|Error|assert" $O/r02k_pytest_fused.log | cut -c1-400 | head -20
for v in 0 1; do
  echo "== SGR_BWD2_FASTEXP=$v"
  SGR_BWD2_FASTEXP=$v timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "golden or live_reference or oracle_small or config_C or config_B or precomp or smoke_script or callsite or edge or sums_to_whole" 2>&1 | tail -4 | cut -c1-300
  SGR_BWD2_FASTEXP=$v timeout 200 python bench.py --no-e2e --no-cpu-baseline 2>$O/r02k_bench_$v.err | tail -1 > $O/r02k_bench_$v.json
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/r02k_bench_$v.json').read()); print('ms/step', round(j['ms_per_step'],4), j['config']['stage_ms'])
except Exception as e: print('fail', e); print(open('gpurun_out/r02k_bench_$v.err').read()[-600:])
PY
  SGR_BWD2_FASTEXP=$v timeout 300 python tools/bench_all.py C C_s0.05 > $O/r02k_bench_all_$v.log 2>&1; grep -oE '^C[_s0-9.]* |"grad_rel_vs_ref": \{[^}]*\}|"speedup_fwdbwd": [0-9.]+|"rgb_maxabs_vs_ref": [0-9.e-]+' $O/r02k_bench_all_$v.log | tr '\n' ' '; echo
done
echo done
