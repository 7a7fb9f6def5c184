set -x
timeout -s KILL 200 python -m pytest tests/test_gpu_preprocess.py -m gpu -x -q -k "polya" 2>&1 | tail -8
