#!/bin/bash
set -u
O=gpurun_out
for sch in 1 0; do
  DAV_MATVEC_SCHEDULE=$sch timeout 100 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:matvec_kernel -c 4 \
    python scripts/matvec_only.py --n 100000 --widths 64,128 > $O/ncu_traffic_sched$sch.log 2>&1
  echo "== schedule $sch" | tee -a $O/traffic_steps.log
  grep "matvec_kernel\|dram__\|gpu__time" $O/ncu_traffic_sched$sch.log | sed 's/(CUtensorMap.*//' | tee -a $O/traffic_steps.log
done
