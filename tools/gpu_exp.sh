#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_voc_gpu.py -m gpu -x -q -k "synthesizer or peak or pcm or generate or golden" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-text2wav --no-streaming > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('ms_per_step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'], d['e2e']['value'])
PY
