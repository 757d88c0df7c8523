set -x
python -m pytest tests/test_ops_gpu.py tests/test_guards_gpu.py -x -q -m gpu -k "roi_align or relation_head or guard" 2>&1 | tail -5
python tools/bench_roi.py 2>&1 | tail -5
