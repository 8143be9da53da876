#!/bin/bash
# Validation + A/B of the tensor-pipe tall-skinny GEMM (DAV_GEMM_IMPL=1): unit test, full suite under it, bench phase times.
set -u
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
step() { echo "=== [$(( $(date +%s) - T0 )) s] $*" | tee -a $O/gemm_steps.log; }
step "gemm unit test (both impls)"
timeout 300 python -m pytest tests -m gpu -x -q -k "gemm_tensor_pipe" > $O/pytest_gemm_unit.log 2>&1; echo "rc=$?" | tee -a $O/gemm_steps.log; tail -3 $O/pytest_gemm_unit.log | tee -a $O/gemm_steps.log
step "full suite under DAV_GEMM_IMPL=1"
DAV_GEMM_IMPL=1 timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gemm1.log 2>&1; echo "rc=$?" | tee -a $O/gemm_steps.log; tail -3 $O/pytest_gemm1.log | tee -a $O/gemm_steps.log
for impl in 0 1; do
  step "bench DAV_GEMM_IMPL=$impl"
  DAV_GEMM_IMPL=$impl timeout 300 python bench.py --no-e2e --no-cpu > $O/bench_gemm$impl.json 2> $O/bench_gemm$impl.err; echo "rc=$?" | tee -a $O/gemm_steps.log
  python - <<PY | tee -a $O/gemm_steps.log
import json
d=json.loads(open("$O/bench_gemm$impl.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "phase_ms", d["phase_ms"])
PY
done
step "bench n=20000 both impls"
for impl in 0 1; do
  DAV_GEMM_IMPL=$impl timeout 300 python bench.py --n 20000 --lowest 10 --max-dim 100 --no-e2e --no-cpu > $O/bench_n20k_gemm$impl.json 2>/dev/null
  python - <<PY | tee -a $O/gemm_steps.log
import json
d=json.loads(open("$O/bench_n20k_gemm$impl.json").read().strip().splitlines()[-1])
print("impl $impl n=20000 ms_per_step", d["ms_per_step"], "phase_ms", d["phase_ms"])
PY
done
step "done"
