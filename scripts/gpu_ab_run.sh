#!/bin/bash
# One GPU-box job: full parity suite, A/B of the block-matvec schedules / stage depths, bench and ncu captures.
# Everything lands in gpurun_out/ (merged back by gpurun).  Steps are ordered by importance; each has its own timeout.
set -u
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
step() { echo "=== [$(( $(date +%s) - T0 )) s] $*" | tee -a $O/steps.log; }
nvidia-smi -L > $O/gpu.txt 2>&1; nproc >> $O/gpu.txt; free -g >> $O/gpu.txt
cp MEASURED_PEAKS.json $O/ 2>/dev/null

step "pytest -m gpu (default schedule)"
timeout 900 python -m pytest tests -m gpu -x -q --durations=12 > $O/pytest_default.log 2>&1; echo "rc=$?" | tee -a $O/steps.log
tail -3 $O/pytest_default.log | tee -a $O/steps.log

step "matvec A/B n=100000"
timeout 300 python scripts/matvec_ab.py --n 100000 --widths 16,32,64 --reps 5 --out $O/matvec_ab_n100k.json > $O/matvec_ab_n100k.log 2>&1; echo "rc=$?" | tee -a $O/steps.log
cat $O/matvec_ab_n100k.log | tee -a $O/steps.log

step "matvec A/B without the L2 evict_last hint on X (schedules 0 and 2)"
DAV_MATVEC_L2_HINTS=0 timeout 200 python scripts/matvec_ab.py --n 100000 --widths 16,32,64 --reps 5 --schedules 0,2 --bks 16 > $O/matvec_ab_n100k_nohint.log 2>&1; echo "rc=$?" | tee -a $O/steps.log
cat $O/matvec_ab_n100k_nohint.log | tee -a $O/steps.log

step "pytest -m gpu under DAV_MATVEC_SCHEDULE=2"
DAV_MATVEC_SCHEDULE=2 timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_sched2.log 2>&1; echo "rc=$?" | tee -a $O/steps.log
tail -3 $O/pytest_sched2.log | tee -a $O/steps.log

step "bench default (full: e2e + cpu baseline)"
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?" | tee -a $O/steps.log

step "bench schedule 2 / bk 16"
DAV_MATVEC_SCHEDULE=2 timeout 300 python bench.py --no-e2e --no-cpu > $O/bench_sched2_bk16.json 2> $O/bench_sched2_bk16.err; echo "rc=$?" | tee -a $O/steps.log
step "bench schedule 2 / bk 32"
DAV_MATVEC_SCHEDULE=2 DAV_MATVEC_BK=32 timeout 300 python bench.py --no-e2e --no-cpu > $O/bench_sched2_bk32.json 2> $O/bench_sched2_bk32.err; echo "rc=$?" | tee -a $O/steps.log
step "bench schedule 1 / bk 16"
DAV_MATVEC_SCHEDULE=1 timeout 300 python bench.py --no-e2e --no-cpu > $O/bench_sched1_bk16.json 2> $O/bench_sched1_bk16.err; echo "rc=$?" | tee -a $O/steps.log

step "pytest subset under schedule 2 + bk 32"
DAV_MATVEC_SCHEDULE=2 DAV_MATVEC_BK=32 timeout 600 python -m pytest tests -m gpu -x -q -k "matvec or dense or config2 or callback" > $O/pytest_sched2_bk32.log 2>&1; echo "rc=$?" | tee -a $O/steps.log
tail -3 $O/pytest_sched2_bk32.log | tee -a $O/steps.log

step "ncu --set full, matvec kernels, schedule 2 (bk 16 and 32)"
for BK in 16 32; do
  DAV_MATVEC_SCHEDULE=2 DAV_MATVEC_BK=$BK timeout 400 ncu --set full --clock-control none --import-source on -k regex:matvec_kernel -c 3 \
    -f -o $O/ncu_matvec_sched2_bk$BK python scripts/matvec_only.py --n 100000 --widths 16,32,64 > $O/ncu_matvec_sched2_bk$BK.log 2>&1
  echo "rc=$?" | tee -a $O/steps.log
done
step "ncu --set full, matvec kernels, schedule 0 bk 16 (previous default, same box)"
DAV_MATVEC_SCHEDULE=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:matvec_kernel -c 3 \
    -f -o $O/ncu_matvec_sched0_bk16 python scripts/matvec_only.py --n 100000 --widths 16,32,64 > $O/ncu_matvec_sched0_bk16.log 2>&1
echo "rc=$?" | tee -a $O/steps.log

step "ncu launch list of bench.py under schedule 2"
DAV_MATVEC_SCHEDULE=2 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --replay-mode application -c 400 --csv \
  --log-file $O/ncu_launches_bench_sched2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/ncu_launches_bench_sched2.log 2>&1
echo "rc=$?" | tee -a $O/steps.log

step "compute-sanitizer memcheck on the small schedule cases"
DAV_MATVEC_SCHEDULE=2 timeout 240 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -x -q \
  -k "schedules and (1000 or 4097 or 3000)" > $O/sanitizer_sched.log 2>&1; echo "rc=$?" | tee -a $O/steps.log
tail -5 $O/sanitizer_sched.log | tee -a $O/steps.log
step "done"
