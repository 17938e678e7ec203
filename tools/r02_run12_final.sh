#!/bin/bash
# round-2 final 1-GPU call: whole -m gpu suite, both bench arms on configs C / B / E, launch list, ncu --set full of every own kernel
# (incl. the multi-GPU-only ones through the world = 1 fused step), sanitizer.  Outputs under gpurun_out/r02z_*.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
L=$O/r02z_pytest_gpu.log; : > $L
for f in tests/test_compose_gpu.py tests/test_losses_gpu.py tests/test_parity_gpu.py; do
  echo "=== $f" >> $L
  timeout 900 python -m pytest $f -m gpu -q >> $L 2>&1; echo "rc=$?" >> $L
done
grep -E "^===|passed|failed|^FAILED|rc=|AssertionError: " $L | cut -c1-200 | head -30
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02z_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/r02z_smoke.log | cut -c1-400
timeout 400 python bench.py > $O/r02z_bench_sgr.json 2> $O/r02z_bench_sgr.err
timeout 400 python bench.py --impl reference > $O/r02z_bench_ref.json 2> $O/r02z_bench_ref.err
for w in B E; do
  timeout 400 python bench.py --workload $w --no-cpu-baseline > $O/r02z_bench_sgr_$w.json 2> $O/r02z_bench_sgr_$w.err
  timeout 400 python bench.py --workload $w --no-cpu-baseline --impl reference --steps 5 --warmup 3 > $O/r02z_bench_ref_$w.json 2> $O/r02z_bench_ref_$w.err
done
python - <<'PY'
import json
for fn in ('r02z_bench_sgr','r02z_bench_ref','r02z_bench_sgr_B','r02z_bench_ref_B','r02z_bench_sgr_E','r02z_bench_ref_E'):
    try:
        j=json.loads(open('gpurun_out/%s.json'%fn).read().strip().split('\n')[-1])
        print(fn, round(j['ms_per_step'],4), 'e2e', {k:j.get('e2e',{}).get(k) for k in ('ms_per_step','timed_region')}, j.get('gpu_launches'), j['config'].get('stage_ms'), j.get('roofline',{}).get('traffic'), j.get('cpu_baseline'))
    except Exception as e: print(fn,'ERR',e); print(open('gpurun_out/%s.err'%fn).read()[-600:])
PY
# launch list (eager loop so that every kernel is an individual launch)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/r02z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --graph off > $O/r02z_launch_bench.log 2>&1
# full captures
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:"blend_fwd_kernel|blend_bwd2_kernel|preprocess_fwd_kernel|preprocess_bwd_tma_kernel|emit_pairs_kernel|emit_big_kernel|tile_ranges_kernel" -s 28 -c 7 -o $O/r02z_prof_main python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --graph off > $O/r02z_ncu_main.log 2>&1
timeout 600 $NCU -k regex:"compose_fwd_kernel|compose_bwd_kernel|compose_pose_finalize_kernel" -s 3 -c 3 -o $O/r02z_prof_compose python bench.py --steps 1 --warmup 3 --no-cpu-baseline --graph off > $O/r02z_ncu_compose.log 2>&1
timeout 600 $NCU -k regex:"preprocess_fwd_kernel|count_compact_kernel|preprocess_bwd_tma_kernel" -s 8 -c 3 -o $O/r02z_prof_fused python tools/fused_world1.py > $O/r02z_ncu_fused.log 2>&1
ls -la $O/r02z*.ncu-rep; tail -2 $O/r02z_ncu_fused.log
bash tools/sanitize.sh 2>&1 | tail -12
for t in memcheck racecheck synccheck; do cp $O/sanitizer_$t.log $O/r02z_sanitizer_$t.log; done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $O/r02z_smi.txt
echo done
