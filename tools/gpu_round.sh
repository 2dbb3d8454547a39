#!/bin/bash
# One GPU visit: parity tests, smoke, bench (both arms), ncu launch lists (vocoder cfg2 + TTE cfg3), DRAM traffic of one forward.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | head -20 > gpurun_out/cpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
# every launch of one vocoder forward (after 1 warm-up forward = 98 launches)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_voc.csv python tools/prof_voc.py 8 1024 1 0 > gpurun_out/ncu_voc.log 2>&1; echo "ncu voc rc=$?"
# every launch of one TTE inference at cfg3 (B=32, 256 chars -> 2048 units)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_tte.csv python tools/prof_tte.py > gpurun_out/ncu_tte.log 2>&1; echo "ncu tte rc=$?"
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
