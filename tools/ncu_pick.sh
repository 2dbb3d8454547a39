#!/bin/bash
# ncu --set full on selected tc_conv launches of the second forward (index within the 96 tensor-core launches of a forward)
# usage: ncu_pick.sh tag idx [idx...]
tag=$1; shift
mkdir -p gpurun_out
for idx in "$@"; do
  s=$((96 + idx))
  timeout 300 ncu --set full --clock-control none -k regex:tc_conv_kernel -s $s -c 1 --csv --page raw \
      python tools/prof_voc.py 8 1024 1 1 > gpurun_out/ncu_${tag}_k${idx}.csv 2> gpurun_out/ncu_${tag}_k${idx}.err
  echo "idx $idx rc=$? $(wc -c < gpurun_out/ncu_${tag}_k${idx}.csv) bytes"
done
