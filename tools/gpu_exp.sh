#!/bin/bash
mkdir -p gpurun_out
for m in 2 3 4 6; do
PB_TC_RING=$m timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_ring$m.csv python tools/prof_voc.py 8 1024 1 0 > gpurun_out/ncu_voc.log 2>&1; echo "ncu rc=$?"
done
