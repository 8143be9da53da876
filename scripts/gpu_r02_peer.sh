#!/bin/bash
# r02, N-GPU job: exchange self-check + latency table, sharded parity check against the oracle, bench at N with the
# peer transport (and, with AB=1, with NCCL).   usage: scripts/gpu_r02_peer.sh [N]
set -u
N=${1:-2}
O=gpurun_out
TAG=${TAG:-v6}
mkdir -p $O
T0=$(date +%s)
step() { echo "=== [$(( $(date +%s) - T0 )) s] $*" | tee -a $O/r02_peer_steps.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
step "collectives check ($N ranks)"
timeout 300 $TR --master-port 29513 tests/dist_collectives_check.py 100000 > $O/r02_collectives_${N}gpu_$TAG.log 2>&1
echo "rc=$?" | tee -a $O/r02_peer_steps.log; grep "^{\|PASSED\|Error\|error\|assert" $O/r02_collectives_${N}gpu_$TAG.log | tail -5 | tee -a $O/r02_peer_steps.log
step "dist_gpu_check ($N ranks)"
timeout 300 $TR --master-port 29511 tests/dist_gpu_check.py > $O/r02_dist_check_${N}gpu_$TAG.log 2>&1
echo "rc=$?" | tee -a $O/r02_peer_steps.log; grep "^ok\|^comm\|PASSED\|Error\|error" $O/r02_dist_check_${N}gpu_$TAG.log | tee -a $O/r02_peer_steps.log
step "bench --gpus $N (peer transport)"
timeout 400 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu --no-other > $O/r02_bench_${N}gpu_peer_$TAG.json 2> $O/r02_bench_${N}gpu_peer_$TAG.err
echo "rc=$?" | tee -a $O/r02_peer_steps.log; python scripts/bench_brief.py $O/r02_bench_${N}gpu_peer_$TAG.json | tee -a $O/r02_peer_steps.log
if [ "${AB:-0}" = "1" ]; then
step "bench --gpus $N (NCCL transport)"
DAV_PEER_COLLECTIVES=0 timeout 400 $TR --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu --no-other > $O/r02_bench_${N}gpu_nccl_$TAG.json 2> $O/r02_bench_${N}gpu_nccl_$TAG.err
echo "rc=$?" | tee -a $O/r02_peer_steps.log; python scripts/bench_brief.py $O/r02_bench_${N}gpu_nccl_$TAG.json | tee -a $O/r02_peer_steps.log
fi
if [ "${FULL:-0}" = "1" ]; then
step "bench --gpus $N (driver form: e2e + other_configs)"
timeout 900 $TR --master-port 29515 bench.py --gpus $N --steps 10 --warmup 3 > $O/r02_bench_${N}gpu_full_$TAG.json 2> $O/r02_bench_${N}gpu_full_$TAG.err
echo "rc=$?" | tee -a $O/r02_peer_steps.log; python scripts/bench_brief.py $O/r02_bench_${N}gpu_full_$TAG.json | tee -a $O/r02_peer_steps.log
fi
step "done"
