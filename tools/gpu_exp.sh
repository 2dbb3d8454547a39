#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('ms_per_step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'])
print(d['latency']); print(d['streaming'])
PY
tail -3 gpurun_out/bench_quick.err
