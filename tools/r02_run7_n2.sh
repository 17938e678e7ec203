#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29520 tools/trace_step.py > $O/r02_trace_n2.log 2>&1
tail -80 $O/trace_n2_rank0.txt
timeout 200 python tools/trace_step.py > $O/r02_trace_n1.log 2>&1
echo done
