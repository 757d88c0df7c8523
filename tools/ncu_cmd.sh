set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong --no-configs > gpurun_out/r2b_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong --no-configs > gpurun_out/r2b_ncu_bench.log 2>&1
python tools/ncu_head.py 64 3 > gpurun_out/r2b_ncu_head_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'correlate_tc|decode_keys|decode_topk|nms_proposals|roi_align|roi_tables|relation_tc|final_detect' -s 16 -c 8 -o gpurun_out/r2b_head python tools/ncu_head.py 64 3 > gpurun_out/r2b_ncu_head.log 2>&1
ls -la gpurun_out/r2b_head.ncu-rep gpurun_out/r2b_launches.csv; tail -3 gpurun_out/r2b_ncu_head.log
