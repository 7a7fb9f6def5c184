set -x
T=r3k
timeout -s KILL 400 python -m pytest tests/test_gpu_shard.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${T}_shard_tests.log
cat gpurun_out/${T}_shard_tests.log
timeout -s KILL 100 python tools/batch_invariance.py 96 2>&1 | tail -4
