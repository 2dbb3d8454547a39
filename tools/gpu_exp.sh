mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/abl_$tag.csv python tools/prof_voc.py 8 1024 1 0 > /dev/null 2>&1; echo "$tag rc=$?"; }
run new PB_X=1
run single PB_TC_PAIR64=0
python tools/ablate_table.py gpurun_out/abl_new.csv gpurun_out/abl_single.csv | sed -n '1p;21,23p;40,62p;80,90p'
timeout 600 python -m pytest tests/test_voc_gpu.py tests/test_switches_gpu.py -m gpu -x -q 2>&1 | tail -4
for v in "PB_X=1" "PB_TC_PAIR64=0"; do env $v python tools/dev_voc_time.py 8 1024 | tail -1; done
