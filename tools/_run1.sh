set -x
python bench.py > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err; echo rc=$?
tail -c 400 gpurun_out/r3_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3_bench_ref.json 2> gpurun_out/r3_bench_ref.err; echo rc=$?
tail -c 300 gpurun_out/r3_bench_ref.json
