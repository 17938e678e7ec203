#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q > $O/r02g_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r02g_pytest_gpu.log
grep -E "passed|failed|^FAILED|rc=|AssertionError: " $O/r02g_pytest_gpu.log | cut -c1-300 | head
timeout 300 python bench.py --no-cpu-baseline > $O/r02g_bench_n1.json 2> $O/r02g_bench_n1.err
timeout 300 $TR --master-port 29531 tools/check_gaussian_sharded.py p2p > $O/r02g_gs_check_n2_p2p.log 2>&1; echo "rc=$?" >> $O/r02g_gs_check_n2_p2p.log
tail -3 $O/r02g_gs_check_n2_p2p.log
timeout 300 $TR --master-port 29532 bench.py --gpus 2 --no-cpu-baseline > $O/r02g_bench_n2.json 2> $O/r02g_bench_n2.err
timeout 300 $TR --master-port 29533 bench.py --gpus 2 --no-cpu-baseline --graph off --no-e2e > $O/r02g_bench_n2_eager.json 2> $O/r02g_bench_n2_eager.err
timeout 300 $TR --master-port 29534 tools/trace_step.py > $O/r02g_trace_n2.log 2>&1
python - <<'PY'
import json
for fn in ('r02g_bench_n1.json','r02g_bench_n2.json','r02g_bench_n2_eager.json'):
    try:
        j=json.loads(open('gpurun_out/'+fn).read().strip().split('\n')[-1])
        print(fn, j['ms_per_step'], j.get('parity_n'), j.get('gpu_launches'), j['config'].get('timed_region'), j['config'].get('graph_note'), j.get('e2e',{}).get('ms_per_step'))
    except Exception as e: print(fn,'ERR',e); print(open('gpurun_out/'+fn.replace('.json','.err')).read()[-1200:])
PY
grep -E "emit|count_tiles|preprocess|peer_barrier|step =" $O/trace_n2_rank0.txt | head -20
echo done
