set -x
timeout -s KILL 300 python -m pytest tests/test_gpu_preprocess.py -m gpu -x -q 2>&1 | tail -5
