set -x
python -m pytest tests/test_conv_gpu.py -x -q -m gpu -k "split" 2>&1 | tail -12
