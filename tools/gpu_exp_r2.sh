#!/bin/bash
# Round-2 experiments, written in round 1 after the GPU budget was spent (none has run on a GPU yet):
#  (1) column-split epilogue groups in tc_conv (TcParams::epi_split; variant build) + 4-sample register tiles in conv_post (PB_POST_TILE4, same variant)
#  (2) PB_TC_TMA_OUT=3: TMA-store epilogue on the CTA-pair layers with k <= 3 (product library, host-side switch)
#  (3) PB_VOC_CHAINS=1: the three ResBlock chains of a stage on three streams (product library, host-side switch)
#  (5) -DPB_TC_GROUPS4: a fourth epilogue warpgroup, registers handed over with setmaxnreg (variant build; with PB_TC_EPI_SPLIT=1 also on the wide layers)
#  (4) -DPB_TC_FAMILY: tc_conv instantiated for the vocoder's layer family (flag tests of the epilogue folded away; variant build)
# HERE, before gpurun (the .so files travel):
#   python -m parrot_tts_b200.build --variant episplit --define PB_TC_EPI_SPLIT --define PB_POST_TILE4
#   python -m parrot_tts_b200.build --variant family --define PB_TC_FAMILY
#   python -m parrot_tts_b200.build --variant groups4 --define PB_TC_GROUPS4 --define PB_TC_EPI_SPLIT
#   python -m parrot_tts_b200.build --variant r2all --define PB_TC_EPI_SPLIT --define PB_POST_TILE4 --define PB_TC_FAMILY --define PB_TC_GROUPS4
# ON THE BOX:                              bash tools/gpu_exp_r2.sh
# PB_TC_EPI_SPLIT: 0 = off (the variant build then behaves like the product), 1 = CTA-pair layers, 2 = every one-CTA-per-SM layer
mkdir -p gpurun_out
# bitwise parity of every prepared switch against the default path (fresh interpreter per case)
PB_TEST_EXPERIMENTS=1 timeout 1800 python -m pytest tests/test_experiments_gpu.py -m gpu -q 2>&1 | tail -3
export PB_LIB_VARIANT=episplit
for m in 1 2; do
  PB_TC_EPI_SPLIT=$m timeout 900 python -m pytest tests/test_voc_gpu.py tests/test_tte_gpu.py -m gpu -x -q > gpurun_out/episplit_pytest_$m.log 2>&1
  echo "mode $m pytest rc=$?"; tail -2 gpurun_out/episplit_pytest_$m.log
done
for m in 0 1 2; do
  echo "== PB_TC_EPI_SPLIT=$m"
  PB_TC_EPI_SPLIT=$m timeout 300 python tools/dev_voc_time.py 8 1024 | tail -1
  PB_TC_EPI_SPLIT=$m timeout 300 python tools/dev_tte_time.py | tail -1
done
echo "== family"; PB_LIB_VARIANT=family timeout 300 python tools/dev_voc_time.py 8 1024 | tail -1
for m in 0 1; do echo "== groups4 PB_TC_EPI_SPLIT=$m"; PB_LIB_VARIANT=groups4 PB_TC_EPI_SPLIT=$m timeout 300 python tools/dev_voc_time.py 8 1024 | tail -1; done
echo "== r2all + chains"; PB_LIB_VARIANT=r2all PB_TC_EPI_SPLIT=1 PB_VOC_CHAINS=1 timeout 300 python tools/dev_voc_time.py 8 1024 | tail -1
# second prepared switch (product library, host-side only): TMA-store epilogue on the CTA-pair layers with k <= 3
unset PB_LIB_VARIANT
for m in 1 3; do echo "== PB_TC_TMA_OUT=$m"; PB_TC_TMA_OUT=$m timeout 300 python tools/dev_voc_time.py 8 1024 | tail -1; done
PB_TC_TMA_OUT=3 timeout 900 python -m pytest tests/test_voc_gpu.py -m gpu -x -q 2>&1 | tail -1
# (3) chain streams: parity (bit-identical by construction: the running sum keeps its order), CUDA-graph capture, timing
for m in 0 1; do echo "== PB_VOC_CHAINS=$m"; PB_VOC_CHAINS=$m timeout 300 python tools/dev_voc_time.py 8 1024 | tail -1; done
# with programmatic dependent launch the NEXT kernel of the same chain becomes resident early and idles in griddepcontrol.wait on the SMs
# another chain's kernel could use: try the chains without it as well
for m in 0 1; do echo "== PB_TC_NO_PDL=1 PB_VOC_CHAINS=$m"; PB_TC_NO_PDL=1 PB_VOC_CHAINS=$m timeout 300 python tools/dev_voc_time.py 8 1024 | tail -1; done
PB_VOC_CHAINS=1 timeout 900 python -m pytest tests/test_voc_gpu.py -m gpu -x -q 2>&1 | tail -1
# the single-utterance path (cfg1) is where independent chains can really run side by side: stages 1-2 are under one wave there
for m in 0 1; do echo "== latency PB_VOC_CHAINS=$m"; PB_VOC_CHAINS=$m timeout 300 python tools/dev_latency.py | tail -3; done
PB_VOC_CHAINS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-text2wav > gpurun_out/bench_chains.json 2> gpurun_out/bench_chains.err; echo "bench rc=$?"
# per-launch view of the winner (compare with profiles/r01_model_vs_measured.txt):
#   PB_TC_EPI_SPLIT=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
#       --log-file gpurun_out/launches_episplit.csv python tools/prof_voc.py 8 1024 1 0 && python tools/roofline_model.py gpurun_out/launches_episplit.csv
