#!/bin/bash
# round-2 GPU call (8 GPUs): N = 8 (full line incl. composed e2e as a graph) and N = 4 after the block-run exchange + N = 8 kernel timeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
run() { # n port extra...
  n=$1; port=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --no-cpu-baseline "$@"
}
run 8 29614 > $O/r02m_scale_n8.json 2> $O/r02m_scale_n8.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 tools/trace_step.py > $O/r02m_trace_n8.log 2>&1
run 4 29613 --no-e2e > $O/r02m_scale_n4.json 2> $O/r02m_scale_n4.err
python - <<'PY'
import json
for fn in ('r02m_scale_n8.json','r02m_scale_n4.json'):
    try:
        j=json.loads(open('gpurun_out/'+fn).read().strip().split('\n')[-1])
        print(fn, round(j['ms_per_step'],4), j.get('parity_n'), j.get('gpu_launches'), (j['config'].get('timed_region') or '')[:12], j['config'].get('graph_note'), j.get('e2e'))
    except Exception as e: print(fn,'ERR',e); print(open('gpurun_out/'+fn.replace('.json','.err')).read()[-800:])
PY
for r in 0 7; do cut -c1-120 $O/trace_n8_rank$r.txt | grep -v "Memset\|ExclusiveSum\|ScanInit\|CompactInit\|Onesweep\|Histogram\|FillFunctor"; done
echo done
