#!/bin/bash
# launch lists of the last forward (vocoder cfg2, TTE cfg3) + ncu --set full of one decoder attention launch and one FFN conv launch
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_voc.csv python tools/prof_voc.py 8 1024 1 0 > gpurun_out/ncu_voc.log 2>&1; echo "ncu voc rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_tte.csv python tools/prof_tte.py > gpurun_out/ncu_tte.log 2>&1; echo "ncu tte rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_tc_kernel -s 13 -c 1 -f -o /tmp/ncu/attn python tools/prof_tte.py > /tmp/ncu/attn.log 2>&1
ncu -i /tmp/ncu/attn.ncu-rep --page raw --csv > gpurun_out/ncu_attn_dec.csv 2>/dev/null
ncu -i /tmp/ncu/attn.ncu-rep --page source --csv > gpurun_out/src_attn_dec.csv 2>/dev/null
echo "attn rc=$? $(wc -c < gpurun_out/ncu_attn_dec.csv)"
