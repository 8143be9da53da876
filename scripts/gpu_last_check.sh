#!/bin/bash
set -u
O=gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q -k "matvec or dense_dropin or config2 or handle_trace or smoke" > $O/pytest_last.log 2>&1; echo "pytest rc=$?" | tee $O/last_steps.log; tail -2 $O/pytest_last.log | tee -a $O/last_steps.log
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -1 | tee -a $O/last_steps.log
timeout 200 python bench.py --no-e2e --no-cpu > $O/bench_last.json 2>/dev/null; echo "bench rc=$?" | tee -a $O/last_steps.log
