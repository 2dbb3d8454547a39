#!/bin/bash
# launch lists of the last forward (vocoder cfg2, TTE cfg3) + ncu --set full of representative launches:
#   tc_conv stage 1 (C=256, k=11), tc_conv stage 3 (C=64, k=7, residual), tc_pair stage 5 (C=16, k=11), decoder attention
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_voc.csv python tools/prof_voc.py 8 1024 1 0 > gpurun_out/ncu_voc.log 2>&1; echo "ncu voc rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_tte.csv python tools/prof_tte.py > gpurun_out/ncu_tte.log 2>&1; echo "ncu tte rc=$?"
cap() {  # tag kernel-regex skip script...
  local tag=$1 re=$2 skip=$3; shift 3
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$re -s $skip -c 1 -f -o /tmp/ncu/$tag "$@" > /tmp/ncu/$tag.log 2>&1
  ncu -i /tmp/ncu/$tag.ncu-rep --page raw --csv > gpurun_out/ncu_full_$tag.csv 2>/dev/null
  ncu -i /tmp/ncu/$tag.ncu-rep --page source --csv > gpurun_out/src_$tag.csv 2>/dev/null
  echo "$tag rc=$? $(wc -c < gpurun_out/ncu_full_$tag.csv)"
}
# tc_conv launches of one vocoder forward (tensor-core path, pairs fused): conv_pre, ups0, 18 x stage1, ups1, 18 x stage2, ups2,
# 18 x stage3, ups3, ups4 = 60 per forward; prof_voc.py runs 1 forward (warm = 0): indices into that forward
cap tc_conv_s1_k11 tc_conv_kernel 14 python tools/prof_voc.py 8 1024 1 0
cap tc_conv_s3_k7c2 tc_conv_kernel 47 python tools/prof_voc.py 8 1024 1 0
cap tc_pair_s5_k11 tc_pair_kernel 15 python tools/prof_voc.py 8 1024 1 0
cap attn_dec attn_tc_kernel 13 python tools/prof_tte.py
