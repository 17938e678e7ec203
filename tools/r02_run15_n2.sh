#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 tools/check_gaussian_sharded.py p2p > $O/r02l_gs_check_n2_p2p.log 2>&1; echo "rc=$?" >> $O/r02l_gs_check_n2_p2p.log
tail -3 $O/r02l_gs_check_n2_p2p.log | cut -c1-300
timeout 300 $TR --master-port 29542 bench.py --gpus 2 --no-cpu-baseline > $O/r02l_bench_n2.json 2> $O/r02l_bench_n2.err
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/r02l_bench_n2.json').read().strip().split('\n')[-1])
    print(round(j['ms_per_step'],4), j.get('parity_n'), j.get('gpu_launches'), j['config'].get('timed_region','')[:12], j['config'].get('graph_note'), j.get('e2e',{}).get('ms_per_step'), j['config'].get('stage_ms'))
except Exception as e: print('ERR',e); print(open('gpurun_out/r02l_bench_n2.err').read()[-1200:])
PY
timeout 300 $TR --master-port 29544 tools/trace_step.py > $O/r02l_trace_n2.log 2>&1
cut -c1-120 $O/trace_n2_rank0.txt | grep -v "Memset\|ExclusiveSum\|ScanInit\|CompactInit"
echo done
