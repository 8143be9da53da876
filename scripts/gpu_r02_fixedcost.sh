#!/bin/bash
# r02: the fixed (non-matvec) cost of the headline solve as one rank of 8 sees it: n = 12,500 on one GPU has the
# same nl x k shapes in every kernel except the block matvec.  Phase table + ncu launch list.
set -u
O=gpurun_out
TAG=${TAG:-v3}
mkdir -p $O
T0=$(date +%s)
step() { echo "=== [$(( $(date +%s) - T0 )) s] $*" | tee -a $O/r02_fixedcost_steps.log; }
step "bench n=12500 $TAG"
timeout 600 python bench.py --n 12500 --steps 20 --warmup 5 --no-e2e --no-cpu > $O/r02_bench_n12500_$TAG.json 2> $O/r02_bench_n12500_$TAG.err
echo "rc=$?"; python scripts/bench_brief.py $O/r02_bench_n12500_$TAG.json | tee -a $O/r02_fixedcost_steps.log
step "ncu launch list n=12500"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_ncu_launches_n12500_$TAG.csv \
  python bench.py --n 12500 --steps 2 --warmup 1 --no-e2e --no-cpu > $O/r02_ncu_n12500_$TAG.log 2>&1
echo "rc=$?"
python scripts/summarize_launches.py $O/r02_ncu_launches_n12500_$TAG.csv > $O/r02_ncu_launches_n12500_${TAG}_summary.txt 2>&1
head -50 $O/r02_ncu_launches_n12500_${TAG}_summary.txt
step "done"
