#!/bin/bash
# 2-GPU job: sharded parity check against the oracle and the strong-scaling bench line at N=2.
set -u
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
step() { echo "=== [$(( $(date +%s) - T0 )) s] $*" | tee -a $O/gpu2_steps.log; }
step "dist_gpu_check (2 ranks)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > $O/dist_check_2gpu.log 2>&1
echo "rc=$?" | tee -a $O/gpu2_steps.log; grep "^ok\|PASSED\|Error\|error" $O/dist_check_2gpu.log | tee -a $O/gpu2_steps.log
step "bench --gpus 2"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err
echo "rc=$?" | tee -a $O/gpu2_steps.log; tail -c 600 $O/bench_2gpu.json | tee -a $O/gpu2_steps.log
step "done"
