set -x
T=r4g
CS="compute-sanitizer --error-exitcode 9"
( timeout -s KILL 300 $CS --tool memcheck python tools/stress_preprocess.py 200 12 2>&1 | tail -4; echo rc=$? ) > gpurun_out/${T}_pre_memcheck.txt
cat gpurun_out/${T}_pre_memcheck.txt
( timeout -s KILL 400 $CS --tool racecheck python tools/stress_preprocess.py 120 13 2>&1 | tail -4; echo rc=$? ) > gpurun_out/${T}_pre_racecheck.txt
cat gpurun_out/${T}_pre_racecheck.txt
( timeout -s KILL 400 $CS --tool memcheck python -m pytest tests/test_gpu_preprocess.py -m gpu -x -q -k "ranges or corner or edge or outlier" 2>&1 | tail -4; echo rc=$? ) > gpurun_out/${T}_pre_tests_memcheck.txt
cat gpurun_out/${T}_pre_tests_memcheck.txt
( timeout -s KILL 300 $CS --tool synccheck python tools/stress_preprocess.py 120 14 2>&1 | tail -4; echo rc=$? ) > gpurun_out/${T}_pre_synccheck.txt
cat gpurun_out/${T}_pre_synccheck.txt
