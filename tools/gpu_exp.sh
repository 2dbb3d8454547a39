#!/bin/bash
# closing extras: the cfg4 pool on one GPU with the final kernels, and an ncu --set full capture of the time-grouped k = 11 launch
mkdir -p gpurun_out /tmp/ncu
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4_n1.err; echo "cfg4 n1 rc=$?"; tail -c 400 gpurun_out/bench_cfg4_n1.json
timeout 300 ncu --set full --import-source on --clock-control none -k regex:tc_pair16 -s 3 -c 1 -f -o /tmp/ncu/pair16 python tools/prof_voc.py 8 1024 1 0 > /tmp/ncu/pair16.log 2>&1
ncu -i /tmp/ncu/pair16.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_tc_pair16_s5_k11.csv 2>/dev/null; echo "full pair16 rc=$? $(wc -c < gpurun_out/r02_ncu_full_tc_pair16_s5_k11.csv)"
