#!/bin/bash
# Final job of the round: full parity suite on the default build, bench lines, split-target A/B, ncu launch list.
set -u
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
step() { echo "=== [$(( $(date +%s) - T0 )) s] $*" | tee -a $O/final_steps.log; }
nvidia-smi -L > $O/gpu.txt 2>&1; nproc >> $O/gpu.txt; free -g >> $O/gpu.txt
step "pytest -m gpu (default build)"
timeout 600 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_final.log 2>&1; echo "rc=$?" | tee -a $O/final_steps.log; tail -3 $O/pytest_final.log | tee -a $O/final_steps.log
step "bench default (full)"
timeout 600 python bench.py > $O/bench_final.json 2> $O/bench_final.err; echo "rc=$?" | tee -a $O/final_steps.log
for tgt in 296 444 592; do
  step "bench split target $tgt"
  DAV_GEMM_SPLIT_TARGET=$tgt timeout 200 python bench.py --no-e2e --no-cpu > $O/bench_tgt$tgt.json 2>/dev/null
  DAV_GEMM_SPLIT_TARGET=$tgt timeout 200 python bench.py --n 20000 --lowest 10 --max-dim 100 --no-e2e --no-cpu > $O/bench_n20k_tgt$tgt.json 2>/dev/null
  python - <<PY | tee -a $O/final_steps.log
import json
for f in ("$O/bench_tgt$tgt.json", "$O/bench_n20k_tgt$tgt.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print("target $tgt", d["config"]["n"], "ms_per_step %.3f" % d["ms_per_step"], {k: round(v,3) for k,v in d["phase_ms"].items()})
PY
done
step "pytest under DAV_GEMM_SPLIT_TARGET=444"
DAV_GEMM_SPLIT_TARGET=444 timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_tgt444.log 2>&1; echo "rc=$?" | tee -a $O/final_steps.log; tail -2 $O/pytest_tgt444.log | tee -a $O/final_steps.log
step "bench configs[3] gev GJD"
timeout 300 python bench.py --n 50000 --lowest 8 --gev --method GJD --no-e2e --no-cpu > $O/bench_n50k_gev_gjd_final.json 2>/dev/null; echo "rc=$?" | tee -a $O/final_steps.log
step "ncu dram traffic of the b=64 matvec"
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:matvec_kernel -c 2 \
  python scripts/matvec_only.py --n 100000 --widths 64 > $O/ncu_matvec_b64_dram.log 2>&1; echo "rc=$?" | tee -a $O/final_steps.log
grep "matvec_kernel\|dram__\|gpu__time" $O/ncu_matvec_b64_dram.log | tee -a $O/final_steps.log
step "ncu launch list of bench.py"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --replay-mode application -c 400 --csv \
  --log-file $O/ncu_launches_bench_final.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/ncu_launches_bench_final.log 2>&1
echo "rc=$?" | tee -a $O/final_steps.log
step "done"
