#!/bin/bash
mkdir -p gpurun_out
for m in 1 3; do
PB_TC_TMA_OUT=$m timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_t$m.csv python tools/prof_voc.py 8 1024 1 0 > gpurun_out/ncu_voc.log 2>&1; echo "ncu rc=$?"
done
