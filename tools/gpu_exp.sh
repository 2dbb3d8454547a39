#!/bin/bash
# one-off GPU experiment bundle: full-size parity vs eager torch + section times of the vocoder forward
mkdir -p gpurun_out
rm -f gpurun_out/eager_bar.json
timeout 900 python -m pytest tests/test_fullsize_vs_eager_gpu.py -m gpu -x -q > gpurun_out/pytest_eager.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_eager.log
cat gpurun_out/eager_bar.json
for a in 0 0; do PB_TC_ABLATE=$a timeout 120 python tools/dev_ablate.py 2>&1 | tail -1; done | tee gpurun_out/ablate.txt
timeout 600 python -m pytest tests/test_voc_gpu.py -m gpu -x -q 2>&1 | tail -2
