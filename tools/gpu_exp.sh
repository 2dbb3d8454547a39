#!/bin/bash
# scratch: look-ahead on the next weight sub-block's barrier in the tc_conv MMA issuer (PB_TC_PEEK) - correctness, per launch, same box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_switches_gpu.py tests/test_voc_gpu.py tests/test_tte_gpu.py -x -q -m gpu -k "look_ahead or golden or ragged or rectangle or other_ or tensor_core" 2>&1 | tail -4
for v in 0 1 0 1; do
  PB_TC_PEEK=$v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/abl_peek${v}_$RANDOM.csv \
      python tools/prof_voc.py 8 1024 1 0 > /dev/null 2>&1; echo "peek $v rc=$?"
done
ls gpurun_out/abl_peek*.csv
python tools/ablate_table.py $(ls gpurun_out/abl_peek0_*.csv | head -1) $(ls gpurun_out/abl_peek1_*.csv | head -1) $(ls gpurun_out/abl_peek0_*.csv | tail -1) $(ls gpurun_out/abl_peek1_*.csv | tail -1) > gpurun_out/peek_table.txt
grep -n "launch\|s1 k\|s2 k11\|s2 k3 d1\|ups\|conv_pre\|total" gpurun_out/peek_table.txt | head -50
for v in 0 1; do PB_TC_PEEK=$v python tools/dev_ab.py 8 1024; done 2>&1 | grep AB
