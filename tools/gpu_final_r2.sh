#!/bin/bash
# round 2 closing visit: the full GPU suite, smoke(), the default bench line + reference arm, refreshed launch lists (traffic.json, model table)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cfg3_final.json 2> gpurun_out/bench_cfg3_final.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_cfg3_final.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_cfg3_final.json 2>&1; echo "ref rc=$?"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 600 ncu --metrics $M --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_voc_cfg2.csv python tools/prof_voc.py 8 1024 1 0 > /dev/null 2>&1; echo "list voc cfg2 rc=$?"
timeout 900 ncu --metrics $M --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_voc_cfg3.csv python tools/prof_voc.py 32 2048 1 0 > /dev/null 2>&1; echo "list voc cfg3 rc=$?"
timeout 600 ncu --metrics $M --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_tte_cfg3.csv python tools/prof_tte.py > /dev/null 2>&1; echo "list tte rc=$?"
mkdir -p /tmp/ncu
timeout 300 ncu --set full --import-source on --clock-control none -k regex:tc_pair16 -s 3 -c 1 -f -o /tmp/ncu/pair16 python tools/prof_voc.py 8 1024 1 0 > /tmp/ncu/pair16.log 2>&1
ncu -i /tmp/ncu/pair16.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_tc_pair16_s5_k11.csv 2>/dev/null; echo "full pair16 rc=$? $(wc -c < gpurun_out/r02_ncu_full_tc_pair16_s5_k11.csv)"
