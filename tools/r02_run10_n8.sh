#!/bin/bash
# round-2 GPU call (8 GPUs): scaling table N = 8, 4, 2, 1 on ONE box + N = 8 kernel timeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
run() { # n port extra...
  n=$1; port=$2; shift 2
  if [ "$n" = "1" ]; then timeout 300 python bench.py --no-cpu-baseline "$@"; else
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --no-cpu-baseline "$@"; fi
}
run 8 29601 > $O/r02_scale_n8.json 2> $O/r02_scale_n8.err
tail -c 400 $O/r02_scale_n8.json; echo
run 4 29602 > $O/r02_scale_n4.json 2> $O/r02_scale_n4.err
run 2 29603 > $O/r02_scale_n2.json 2> $O/r02_scale_n2.err
run 1 29604 > $O/r02_scale_n1.json 2> $O/r02_scale_n1.err
run 8 29605 --graph off --no-e2e > $O/r02_scale_n8_eager.json 2> $O/r02_scale_n8_eager.err
run 8 29606 --mp-mode gaussian-p2p-allgather --no-e2e > $O/r02_scale_n8_literal.json 2> $O/r02_scale_n8_literal.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29607 tools/trace_step.py > $O/r02_trace_n8.log 2>&1
python - <<'PY'
import json
for fn in ('r02_scale_n1.json','r02_scale_n2.json','r02_scale_n4.json','r02_scale_n8.json','r02_scale_n8_eager.json','r02_scale_n8_literal.json'):
    try:
        j=json.loads(open('gpurun_out/'+fn).read().strip().split('\n')[-1])
        print(fn, round(j['ms_per_step'],4), j.get('parity_n'), j.get('gpu_launches'), (j['config'].get('timed_region') or '')[:12], j['config'].get('graph_note'), j.get('e2e',{}).get('ms_per_step'))
    except Exception as e: print(fn,'ERR',e); print(open('gpurun_out/'+fn.replace('.json','.err')).read()[-800:])
PY
cat $O/trace_n8_rank0.txt | cut -c1-120
echo done
