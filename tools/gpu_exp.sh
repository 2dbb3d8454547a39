#!/bin/bash
# scratch: residual / running-sum requests issued after the TMA-store fence (variant reqlate) vs before it (product), same box
mkdir -p gpurun_out
for v in "" reqlate "" reqlate; do PB_LIB_VARIANT=$v python tools/dev_ab.py 8 1024; done 2>&1 | grep AB
for v in base reqlate; do
  PB_LIB_VARIANT=$( [ $v = base ] && echo "" || echo $v ) timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -c 400 --csv --log-file gpurun_out/abl_$v.csv \
      python tools/prof_voc.py 8 1024 1 0 > /dev/null 2>&1; echo "$v rc=$?"
done
python tools/ablate_table.py gpurun_out/abl_base.csv gpurun_out/abl_reqlate.csv > gpurun_out/reqlate_table.txt
grep -n "s3\|s4\|total" gpurun_out/reqlate_table.txt
