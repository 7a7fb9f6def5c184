set -x
T=r3a
timeout -s KILL 600 python -m pytest tests/test_gpu_preprocess.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
for seed in 2 3; do
timeout -s KILL 300 python tools/stress_preprocess.py 1500 $seed 2>&1 | grep -v "^frame" | tail -6 >> gpurun_out/${T}_stress.log
done
cat gpurun_out/${T}_stress.log
timeout -s KILL 120 python tools/time_preprocess.py 2>/dev/null | tail -3 > gpurun_out/${T}_pre.log
cat gpurun_out/${T}_pre.log
