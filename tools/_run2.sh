set -x
ncu --set full --clock-control none --import-source on -k regex:'roi_' -s 3 -c 3 -o gpurun_out/r3_roi_final python tools/bench_roi.py 64 once > gpurun_out/r3_ncu_roi.log 2>&1
tail -2 gpurun_out/r3_ncu_roi.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong --no-configs > gpurun_out/r3_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r3_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong --no-configs > gpurun_out/r3_ncu_bench.log 2>&1
ls -la gpurun_out/
