set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong --no-configs > gpurun_out/r2c_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong --no-configs > gpurun_out/r2c_ncu_bench.log 2>&1
python tools/ncu_stem1.py > gpurun_out/r2c_stem_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:stem1_tc -s 2 -c 1 -o gpurun_out/r2c_stem1 python tools/ncu_stem1.py > gpurun_out/r2c_ncu_stem.log 2>&1
ls -la gpurun_out/r2c_stem1.ncu-rep gpurun_out/r2c_launches.csv; tail -2 gpurun_out/r2c_ncu_stem.log
