#!/bin/bash
# round 2 profiles: launch lists (time + DRAM bytes per launch) of the cfg2 / cfg3 vocoder forward and the cfg3 TTE forward, ncu --set full of
# the HBM-bound kernels the north star lists (regulator, LayerNorm, conv_post, embedding, ConvTranspose launches, fused head GEMM, mel),
# and compute-sanitizer (memcheck + racecheck) on the default path.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out /tmp/ncu
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 600 ncu --metrics $M --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_voc_cfg2.csv python tools/prof_voc.py 8 1024 1 0 > /dev/null 2>&1; echo "list voc cfg2 rc=$?"
timeout 900 ncu --metrics $M --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_voc_cfg3.csv python tools/prof_voc.py 32 2048 1 0 > /dev/null 2>&1; echo "list voc cfg3 rc=$?"
timeout 600 ncu --metrics $M --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_tte_cfg3.csv python tools/prof_tte.py > /dev/null 2>&1; echo "list tte rc=$?"
cap() {  # tag kernel-regex skip script...
  local tag=$1 re=$2 skip=$3; shift 3
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$re -s $skip -c 1 -f -o /tmp/ncu/$tag "$@" > /tmp/ncu/$tag.log 2>&1
  ncu -i /tmp/ncu/$tag.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_$tag.csv 2>/dev/null
  echo "$tag rc=$? $(wc -c < gpurun_out/r02_ncu_full_$tag.csv)"
}
cap length_regulate length_regulate_kernel 1 python tools/prof_tte.py
cap layernorm_split layernorm_split_kernel 20 python tools/prof_tte.py
cap head_gemm_argmax tc_conv_kernel 65 python tools/prof_tte.py
cap embed_pos embed_pos_kernel 1 python tools/prof_tte.py
cap conv_post conv_post_split_kernel 0 python tools/prof_voc.py 8 1024 1 0
cap voc_embed voc_embed_split_kernel 0 python tools/prof_voc.py 8 1024 1 0
cap ups2_convtranspose tc_conv_kernel 39 python tools/prof_voc.py 8 1024 1 0
cap ups4_convtranspose tc_conv_kernel 59 python tools/prof_voc.py 8 1024 1 0
cap tc_pair_s4_k3 tc_pair_kernel 0 python tools/prof_voc.py 8 1024 1 0
cap mel mel_spectrogram_kernel 1 python tools/prof_mel.py
python tools/ncu_summ.py gpurun_out/r02_ncu_full_*.csv > gpurun_out/r02_ncu_summary.txt 2>&1
bash tools/gpu_sanitize.sh
