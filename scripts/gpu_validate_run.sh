#!/bin/bash
# Validation job of the default build: full parity suite, matvec stress over every schedule / stage depth, bench lines
# of the BASELINE configs, ncu launch list + full capture of the matvec kernels.  Output in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
step() { echo "=== [$(( $(date +%s) - T0 )) s] $*" | tee -a $O/validate_steps.log; }
nvidia-smi -L > $O/gpu.txt 2>&1; nproc >> $O/gpu.txt; free -g >> $O/gpu.txt
step "pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q --durations=12 > $O/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $O/validate_steps.log
tail -4 $O/pytest_gpu.log | tee -a $O/validate_steps.log
step "matvec stress"
timeout 300 python scripts/matvec_stress.py --variants 0:32,1:32,2:32,0:16,1:16,2:16 --reps 40 > $O/stress_final.log 2>&1; echo "rc=$?" | tee -a $O/validate_steps.log
cat $O/stress_final.log | cut -c1-160 | tee -a $O/validate_steps.log
step "smoke"
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?" | tee -a $O/validate_steps.log; tail -2 $O/smoke.log | tee -a $O/validate_steps.log
step "bench default (full)"
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?" | tee -a $O/validate_steps.log
step "bench --impl reference"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc=$?" | tee -a $O/validate_steps.log
step "bench configs[1] n=20000"
timeout 300 python bench.py --n 20000 --lowest 10 --max-dim 100 --no-e2e --no-cpu > $O/bench_n20k.json 2> $O/bench_n20k.err; echo "rc=$?" | tee -a $O/validate_steps.log
step "bench configs[3] n=50000 gev GJD"
timeout 300 python bench.py --n 50000 --lowest 8 --gev --method GJD --no-e2e --no-cpu > $O/bench_n50k_gev_gjd.json 2> $O/bench_n50k_gev_gjd.err; echo "rc=$?" | tee -a $O/validate_steps.log
step "matvec A/B (final build)"
timeout 300 python scripts/matvec_ab.py --n 100000 --widths 8,16,32,64,128 --reps 5 --schedules 0,1 --bks 16,32 --out $O/matvec_ab_final.json > $O/matvec_ab_final.log 2>&1; echo "rc=$?" | tee -a $O/validate_steps.log
cat $O/matvec_ab_final.log | tee -a $O/validate_steps.log
step "ncu launch list of bench.py"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --replay-mode application -c 400 --csv \
  --log-file $O/ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/ncu_launches_bench.log 2>&1
echo "rc=$?" | tee -a $O/validate_steps.log
step "ncu --set full of the matvec kernels (default build: bk 32, stream-K)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:matvec_kernel -c 3 \
    -f -o $O/ncu_matvec_final python scripts/matvec_only.py --n 100000 --widths 16,32,64 > $O/ncu_matvec_final.log 2>&1
echo "rc=$?" | tee -a $O/validate_steps.log
step "done"
