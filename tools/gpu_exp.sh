#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_voc_gpu.py -m gpu -x -q -k "synthesizer or peak or pcm or stream" 2>&1 | tail -15
