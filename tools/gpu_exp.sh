mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/abl_$tag.csv python tools/prof_voc.py 8 1024 1 0 > /dev/null 2>&1; echo "$tag rc=$?"; }
run base PB_X=1
run k1_3 PB_TC_PAIR64_MINK1=3
run k2_3 PB_TC_PAIR64_MINK2=3
run k2_3_tma PB_TC_PAIR64_MINK2=3 PB_TC_TMA_OUT=2
python tools/ablate_table.py gpurun_out/abl_base.csv gpurun_out/abl_k1_3.csv gpurun_out/abl_k2_3.csv gpurun_out/abl_k2_3_tma.csv | sed -n '1p;41,53p;81,82p'
