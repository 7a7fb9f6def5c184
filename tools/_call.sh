set -x
T=r3m
timeout -s KILL 200 python -m pytest tests/test_gpu_resnet.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
grep -q failed gpurun_out/${T}_tests.log && exit 1
for rep in 1 2; do
timeout -s KILL 100 python tools/time_resnet.py 512 12048 basic 2>/dev/null | tail -1 >> gpurun_out/${T}_resnet.log
done
timeout -s KILL 100 python tools/time_resnet.py 4096 12048 basic 2>/dev/null | tail -1 >> gpurun_out/${T}_resnet.log
timeout -s KILL 100 python tools/time_resnet.py 512 12048 bottleneck 2>/dev/null | tail -1 >> gpurun_out/${T}_resnet.log
cat gpurun_out/${T}_resnet.log
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_resnet_launches.csv python tools/time_resnet.py 512 12048 basic > /dev/null 2>&1
python tools/summarise_ncu.py launches gpurun_out/${T}_resnet_launches.csv | head -8
