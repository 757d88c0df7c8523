set -x
python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "roi_align or relation_head" 2>&1 | tail -5
python tools/bench_roi.py 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong --no-configs > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 600 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['head']['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['head']['kernels'].items()}, d['clocks'])
PY
