#!/bin/bash
# scratch bundle for the one-off GPU experiment of the moment (rewritten as needed).  Template: per-launch durations of one 8 x 1024-unit
# forward under two settings of a PB_* switch, same box:
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/gpu_exp.sh PB_TP_TOEPLITZ 0 1'
mkdir -p gpurun_out
sw=${1:-PB_TP_TOEPLITZ}; shift
vals=${@:-0 1}
files=""
for v in $vals; do
  env $sw=$v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/abl_${sw}_$v.csv \
      python tools/prof_voc.py 8 1024 1 0 > /dev/null 2>&1; echo "$sw=$v rc=$?"
  files="$files gpurun_out/abl_${sw}_$v.csv"
done
python tools/ablate_table.py $files | tee gpurun_out/exp_table.txt | tail -40
for v in $vals; do env $sw=$v python tools/dev_ab.py 8 1024; done 2>&1 | grep AB
