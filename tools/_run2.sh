set -x
ncu --set full --clock-control none --import-source on -k regex:'roi_align_tiles' -s 1 -c 1 -o gpurun_out/r3_roi_v3 python tools/bench_roi.py 64 once > gpurun_out/r3_ncu_roi.log 2>&1
tail -3 gpurun_out/r3_ncu_roi.log
