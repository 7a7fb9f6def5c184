set -x
T=r4a
timeout -s KILL 300 python -m pytest tests/test_gpu_preprocess.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
grep -q failed gpurun_out/${T}_tests.log && exit 1
timeout -s KILL 200 python tools/stress_preprocess.py 1500 9 2>&1 | grep -v "^frame" | tail -3
timeout -s KILL 100 python tools/time_normalise.py 2>/dev/null | tail -1 | cut -c1-120
timeout -s KILL 100 python tools/time_preprocess.py 2>/dev/null | tail -2
