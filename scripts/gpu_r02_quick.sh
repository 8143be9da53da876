#!/bin/bash
# r02 iteration loop: selected GPU tests + the solve at the 8-GPU shard shape (n = 12,500) and at full size.
set -u
O=gpurun_out
TAG=${TAG:-v4}
mkdir -p $O
T0=$(date +%s)
step() { echo "=== [$(( $(date +%s) - T0 )) s] $*" | tee -a $O/r02_quick_steps.log; }
step "pytest ${PYTEST_K:-gemm or dense or lowest16}"
timeout 1200 python -m pytest tests -x -q -m gpu -k "${PYTEST_K:-gemm or dense or lowest16}" > $O/r02_quick_pytest_$TAG.log 2>&1
echo "rc=$?" | tee -a $O/r02_quick_steps.log; tail -4 $O/r02_quick_pytest_$TAG.log | tee -a $O/r02_quick_steps.log
step "bench n=12500 $TAG"
timeout 600 python bench.py --n 12500 --steps 20 --warmup 5 --no-e2e --no-cpu > $O/r02_bench_n12500_$TAG.json 2> $O/r02_bench_n12500_$TAG.err
echo "rc=$?"; python scripts/bench_brief.py $O/r02_bench_n12500_$TAG.json | tee -a $O/r02_quick_steps.log
step "bench n=100000 $TAG"
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-other > $O/r02_bench_n100k_$TAG.json 2> $O/r02_bench_n100k_$TAG.err
echo "rc=$?"; python scripts/bench_brief.py $O/r02_bench_n100k_$TAG.json | tee -a $O/r02_quick_steps.log
tail -3 $O/r02_bench_n100k_$TAG.err
step "done"
