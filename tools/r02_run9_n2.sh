#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "sharded or fused or capacity" > $O/r02h_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r02h_pytest_gpu.log
grep -E "passed|failed|^FAILED|rc=|AssertionError: " $O/r02h_pytest_gpu.log | cut -c1-300 | head
timeout 300 $TR --master-port 29541 tools/check_gaussian_sharded.py p2p > $O/r02h_gs_check_n2_p2p.log 2>&1; echo "rc=$?" >> $O/r02h_gs_check_n2_p2p.log
tail -2 $O/r02h_gs_check_n2_p2p.log
timeout 300 $TR --master-port 29542 bench.py --gpus 2 --no-cpu-baseline --no-e2e > $O/r02h_bench_n2.json 2> $O/r02h_bench_n2.err
tail -c 700 $O/r02h_bench_n2.json; echo
timeout 300 $TR --master-port 29544 tools/trace_step.py > $O/r02h_trace_n2.log 2>&1
grep -E "emit|count_tiles|preprocess|peer_barrier|step =|idle" $O/trace_n2_rank0.txt | head -20
echo done
