set -x
python -m pytest tests/test_model_gpu.py tests/test_conv_gpu.py -x -q -m gpu 2>&1 | tail -8
python tools/dbg_backbone.py 2>&1 | tail -12
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong --no-configs > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 600 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['head']['ms_per_step'])
for l in d['feature_extractor']['conv2d_nhwc']['layers'][:8]: print(l)
PY
