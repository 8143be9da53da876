#!/bin/bash
# r02: whole GPU suite + the solve at the 8-GPU shard shape and at full size
set -u
O=gpurun_out; TAG=${TAG:-v6}; mkdir -p $O
T0=$(date +%s)
step() { echo "=== [$(( $(date +%s) - T0 )) s] $*" | tee -a $O/r02_full_steps.log; }
step "eigh bench"
python scripts/eigh_bench.py ${KS:-6,12,16,20,24,32,48,64,128} 2>&1 | grep "^k" | tee $O/r02_eigh_bench_$TAG.txt
step "pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02_pytest_gpu_$TAG.log 2>&1
echo "rc=$?" | tee -a $O/r02_full_steps.log; tail -5 $O/r02_pytest_gpu_$TAG.log | tee -a $O/r02_full_steps.log
step "bench n=12500"
timeout 600 python bench.py --n 12500 --steps 20 --warmup 5 --no-e2e --no-cpu > $O/r02_bench_n12500_$TAG.json 2> $O/r02_bench_n12500_$TAG.err
python scripts/bench_brief.py $O/r02_bench_n12500_$TAG.json | tee -a $O/r02_full_steps.log
step "bench n=100000"
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $O/r02_bench_n100k_$TAG.json 2> $O/r02_bench_n100k_$TAG.err
echo "rc=$?"; python scripts/bench_brief.py $O/r02_bench_n100k_$TAG.json | tee -a $O/r02_full_steps.log
tail -3 $O/r02_bench_n100k_$TAG.err
step "done"
