#!/bin/bash
# r02, 1-GPU job: GPU test suite, headline bench, (optional) full-size oracle golden on the box's host cores.
set -u
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
step() { echo "=== [$(( $(date +%s) - T0 )) s] $*" | tee -a $O/r02_single_steps.log; }
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader | head -2 | tee -a $O/r02_single_steps.log
free -g | head -2 | tee -a $O/r02_single_steps.log; nproc | tee -a $O/r02_single_steps.log
step "pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02_pytest_gpu.log 2>&1
echo "rc=$?" | tee -a $O/r02_single_steps.log; tail -5 $O/r02_pytest_gpu.log | tee -a $O/r02_single_steps.log
step "bench (default)"
timeout 900 python bench.py ${BENCH_ARGS:-} > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err
echo "rc=$?" | tee -a $O/r02_single_steps.log; python scripts/bench_brief.py $O/r02_bench_1gpu.json | tee -a $O/r02_single_steps.log
if [ "${GOLDEN:-0}" = "1" ]; then
  step "oracle golden n=100k (host cores)"
  timeout 1200 python tests/golden/make_golden_n100k.py $O/config2_n100k_oracle.json > $O/r02_golden_n100k.log 2>&1
  echo "rc=$?" | tee -a $O/r02_single_steps.log; tail -2 $O/r02_golden_n100k.log | tee -a $O/r02_single_steps.log
fi
step "done"
