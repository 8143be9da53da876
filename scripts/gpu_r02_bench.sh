#!/bin/bash
# r02, 1-GPU job: headline bench with the attached other_configs and the parity assertions.
set -u
O=gpurun_out
TAG=${TAG:-v3}
mkdir -p $O
T0=$(date +%s)
step() { echo "=== [$(( $(date +%s) - T0 )) s] $*" | tee -a $O/r02_bench_steps.log; }
step "bench (default) $TAG"
timeout 1200 python bench.py ${BENCH_ARGS:-} > $O/r02_bench_1gpu_$TAG.json 2> $O/r02_bench_1gpu_$TAG.err
echo "rc=$?" | tee -a $O/r02_bench_steps.log; python scripts/bench_brief.py $O/r02_bench_1gpu_$TAG.json | tee -a $O/r02_bench_steps.log
tail -5 $O/r02_bench_1gpu_$TAG.err
step "done"
