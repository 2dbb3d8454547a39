#!/bin/bash
# quick GPU check: vocoder tests + plan dump + timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_voc_gpu.py -m gpu -x -q > gpurun_out/pytest_voc.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_voc.log
PB_TC_DEBUG=1 timeout 120 python tools/prof_voc.py 8 1024 1 0 2> gpurun_out/plan.txt | tail -1
sort -u gpurun_out/plan.txt > gpurun_out/plan_u.txt; wc -l gpurun_out/plan_u.txt
timeout 300 python tools/dev_voc_time.py 8 1024 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-text2wav > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('ms_per_step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'])
print({k:round(v,3) for k,v in d['roofline']['section_ms'].items()})
PY
