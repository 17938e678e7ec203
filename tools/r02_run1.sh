#!/bin/bash
# round-2 GPU call 1 (1 GPU): tests, both bench arms, config sweep, knn timing, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu.log
tail -5 $O/r02_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02_smoke.log
tail -4 $O/r02_smoke.log
timeout 300 python bench.py --impl reference > $O/r02_bench_ref.json 2> $O/r02_bench_ref.err
timeout 300 python bench.py > $O/r02_bench_sgr.json 2> $O/r02_bench_sgr.err
cat $O/r02_bench_sgr.json | head -c 1500; echo
timeout 900 python tools/bench_all.py C C_sh1 C_s0.02 C_s0.05 B > $O/r02_bench_all.log 2>&1; cp $O/bench_all.json $O/r02_bench_all.json 2>/dev/null
timeout 300 python tools/knn_bench.py > $O/r02_knn.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/r02_launch_bench.log 2>&1
echo done
