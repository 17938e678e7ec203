#!/bin/bash
# round-2 GPU call 4 (2 GPUs): Gaussian-sharded correctness (staged + fused) and bench at N = 2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/check_gaussian_sharded.py p2p > $O/r02_gs_check_n2_p2p.log 2>&1; echo "rc=$?" >> $O/r02_gs_check_n2_p2p.log
tail -8 $O/r02_gs_check_n2_p2p.log
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --diag > $O/r02_bench_n2.json 2> $O/r02_bench_n2.err
tail -c 1800 $O/r02_bench_n2.json; echo; grep diag $O/r02_bench_n2.err | head -4; tail -5 $O/r02_bench_n2.err
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --mp-mode gaussian-p2p-staged --no-e2e --no-cpu-baseline > $O/r02_bench_n2_staged.json 2> $O/r02_bench_n2_staged.err
tail -c 600 $O/r02_bench_n2_staged.json; echo
timeout 300 $TR --master-port 29514 bench.py --gpus 2 --mp-mode gaussian-p2p-allgather --no-e2e --no-cpu-baseline > $O/r02_bench_n2_literal.json 2> $O/r02_bench_n2_literal.err
tail -c 600 $O/r02_bench_n2_literal.json; echo
echo done
