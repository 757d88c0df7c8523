set -x
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1400 -c 340 --csv --log-file gpurun_out/r3b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong --no-configs > gpurun_out/r3b_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 36 -c 2 -o gpurun_out/r3_conv_split python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong --no-configs > gpurun_out/r3_ncu_conv.log 2>&1
ls -la gpurun_out | tail -8
