#!/bin/bash
# per-launch durations of one 8 x 1024-unit vocoder forward under PB_TC_ABLATE settings (dev switch of tc_conv.cu: 1 = no MMAs,
# 2 = no output stores, 4 = no residual / running-sum loads, 8 = no activation TMA loads, 16 = no accumulator reads, 32 = no epilogue arithmetic)
mkdir -p gpurun_out
for a in "$@"; do
  PB_TC_ABLATE=$a timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/abl_$a.csv \
      python tools/prof_voc.py 8 1024 1 0 > /dev/null 2>&1; echo "ablate $a rc=$?"
done
python tools/ablate_table.py $(for a in "$@"; do echo gpurun_out/abl_$a.csv; done) > gpurun_out/ablate_table.txt
