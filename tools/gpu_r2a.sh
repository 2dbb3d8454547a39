#!/bin/bash
# round 2, call A: full GPU suite, the new default bench line (cfg3 text -> wav), the reference arm, the two remaining switch timings,
# and source-level ncu captures of two epilogue-bound launches (stage-1 / stage-3 k = 3 residual convs)
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_cfg3.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_cfg3.json 2>&1; echo "ref rc=$?"
for m in 1 3; do echo "== PB_TC_TMA_OUT=$m"; PB_TC_TMA_OUT=$m timeout 300 python tools/dev_voc_time.py 8 1024 | tail -1; done
for m in 0 1; do echo "== latency PB_VOC_CHAINS=$m"; PB_VOC_CHAINS=$m timeout 300 python tools/dev_latency.py | tail -6; done
cap() {  # tag kernel-regex skip script...
  local tag=$1 re=$2 skip=$3; shift 3
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$re -s $skip -c 1 -f -o /tmp/ncu/$tag "$@" > /tmp/ncu/$tag.log 2>&1
  ncu -i /tmp/ncu/$tag.ncu-rep --page raw --csv > gpurun_out/ncu_full_$tag.csv 2>/dev/null
  ncu -i /tmp/ncu/$tag.ncu-rep --page source --csv > gpurun_out/src_$tag.csv 2>/dev/null
  echo "$tag rc=$? $(wc -c < gpurun_out/ncu_full_$tag.csv)"
  python tools/ncu_hot.py gpurun_out/src_$tag.csv 60 > gpurun_out/hot_$tag.txt 2>&1
}
cap tc_conv_s1_k3c2 tc_conv_kernel 3 python tools/prof_voc.py 8 1024 1 0
cap tc_conv_s3_k3c2 tc_conv_kernel 41 python tools/prof_voc.py 8 1024 1 0
