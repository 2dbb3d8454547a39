#!/bin/bash
# per-instruction source page (SASS + stall samples) of selected tc_conv launches; only CSVs are kept
tag=$1; shift
mkdir -p gpurun_out /tmp/ncu
for idx in "$@"; do
  s=$((96 + idx))
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:tc_conv_kernel -s $s -c 1 -f -o /tmp/ncu/r_$idx \
      python tools/prof_voc.py 8 1024 1 1 > /tmp/ncu/log_$idx.txt 2>&1
  ncu -i /tmp/ncu/r_$idx.ncu-rep --page source --csv > gpurun_out/src_${tag}_k${idx}.csv 2>/dev/null
  ncu -i /tmp/ncu/r_$idx.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_k${idx}.csv 2>/dev/null
  echo "idx $idx rc=$? $(wc -c < gpurun_out/src_${tag}_k${idx}.csv) bytes"
done
