#!/bin/bash
# scratch: does backing off in mbarrier waits (fewer shared-memory polls while the tensor core streams its operands) help? same box
mkdir -p gpurun_out
for v in base sleep32 sleep200; do
  PB_LIB_VARIANT=$( [ $v = base ] && echo "" || echo $v ) timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/abl_$v.csv \
      python tools/prof_voc.py 8 1024 1 0 > /dev/null 2>&1; echo "$v rc=$?"
done
python tools/ablate_table.py gpurun_out/abl_base.csv gpurun_out/abl_sleep32.csv gpurun_out/abl_sleep200.csv > gpurun_out/sleep_table.txt
grep -n "launch\|s1 k3 d1\|s1 k11 d1\|s2 k7 d1\|s3 k\|s4 k\|s5 k\|total" gpurun_out/sleep_table.txt
for v in "" sleep32 sleep200; do PB_LIB_VARIANT=$v python tools/dev_ab.py 8 1024; done 2>&1 | grep AB
