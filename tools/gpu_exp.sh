#!/bin/bash
mkdir -p gpurun_out
for m in 0 1; do
PB_TC_GROUPS4=$m timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_g4_$m.csv python tools/prof_voc.py 8 1024 1 0 > gpurun_out/ncu_voc.log 2>&1; echo "ncu rc=$?"
done
PB_TC_GROUPS4=1 timeout 600 python -m pytest tests/test_voc_gpu.py -m gpu -x -q 2>&1 | tail -2
