#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('ms_per_step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'])
print(d['latency']); print(d['streaming']); print(d['text2wav'])
PY
tail -5 gpurun_out/bench_quick.err
