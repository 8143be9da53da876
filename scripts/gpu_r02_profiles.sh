#!/bin/bash
# r02 evidence job (1 GPU): driver-form bench line, ncu launch list of the same command, ncu full captures of the
# dominant kernel (b = 64 block matvec) and of the new small kernels, DRAM traffic of the matvec per width.
set -u
O=gpurun_out; TAG=${TAG:-v12}; mkdir -p $O
T0=$(date +%s)
step() { echo "=== [$(( $(date +%s) - T0 )) s] $*" | tee -a $O/r02_profiles_steps.log; }
nvidia-smi -L > $O/r02_gpu.txt 2>&1; nproc >> $O/r02_gpu.txt; free -g >> $O/r02_gpu.txt
step "bench (driver form: default flags)"
timeout 900 python bench.py > $O/r02_bench_1gpu_$TAG.json 2> $O/r02_bench_1gpu_$TAG.err
echo "rc=$?" | tee -a $O/r02_profiles_steps.log; python scripts/bench_brief.py $O/r02_bench_1gpu_$TAG.json | tee -a $O/r02_profiles_steps.log
step "ncu launch list of bench.py"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --replay-mode application -c 500 --csv \
  --log-file $O/r02_ncu_launches_bench_n100k_$TAG.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-other > $O/r02_ncu_launches_bench_$TAG.log 2>&1
echo "rc=$?" | tee -a $O/r02_profiles_steps.log
python scripts/summarize_launches.py $O/r02_ncu_launches_bench_n100k_$TAG.csv "ncu --metrics gpu__time_duration.sum --clock-control none --replay-mode application -c 500 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-other (r02 $TAG build)" "cold-cache, serialised: compare SHARES" > $O/r02_ncu_launches_bench_n100k_${TAG}_summary.txt 2>&1
head -30 $O/r02_ncu_launches_bench_n100k_${TAG}_summary.txt
step "ncu dram traffic of the matvec, b = 16 32 64"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:matvec_kernel -c 3 \
  python scripts/matvec_only.py --n 100000 --widths 16,32,64 > $O/r02_ncu_matvec_dram_$TAG.log 2>&1; echo "rc=$?" | tee -a $O/r02_profiles_steps.log
grep "matvec_kernel\|dram__\|gpu__time" $O/r02_ncu_matvec_dram_$TAG.log | sed 's/(CUtensorMap.*//' | tee $O/r02_ncu_matvec_dram_$TAG.txt
step "ncu full: b = 64 matvec"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:matvec_kernel -c 1 -o $O/r02_ncu_full_matvec_b64_$TAG -f \
  python scripts/matvec_only.py --n 100000 --widths 64 > $O/r02_ncu_full_matvec_b64_$TAG.log 2>&1; echo "rc=$?" | tee -a $O/r02_profiles_steps.log
step "ncu full: the small kernels of one solve at the 8-GPU shard shape (n = 12500), text summary only"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:resid_dmma|pip_small|tridiag_reg|tri_eigvec|guard_cols|gemm_dmma|ll_reduce' -s 20 -c 24 \
  -o /tmp/r02_small -f python bench.py --n 12500 --steps 1 --warmup 1 --no-e2e --no-cpu > $O/r02_ncu_full_small_$TAG.log 2>&1; echo "rc=$?" | tee -a $O/r02_profiles_steps.log
ncu -i /tmp/r02_small.ncu-rep --page raw --csv > /tmp/r02_small.csv 2>/dev/null
python scripts/ncu_pick.py /tmp/r02_small.csv > $O/r02_ncu_full_small_kernels_n12500_$TAG.txt 2>&1
ncu -i $O/r02_ncu_full_matvec_b64_$TAG.ncu-rep --page raw --csv > /tmp/r02_mv.csv 2>/dev/null
python scripts/ncu_pick.py /tmp/r02_mv.csv > $O/r02_ncu_full_matvec_b64_$TAG.txt 2>&1
rm -f $O/r02_ncu_full_matvec_b64_$TAG.ncu-rep
step "done"
