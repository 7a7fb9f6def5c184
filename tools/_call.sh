set -x
timeout 600 python -m pytest tests/test_gpu_resnet.py -m gpu -q -s 2>&1 | tail -15 > gpurun_out/r2f_resnet_tests.log
timeout 120 python tools/debug_resnet_tc.py basic > gpurun_out/r2f_dbg_basic.log 2>&1
for cfg in basic bottleneck; do for tc in 1 0; do RISER_RESNET_TC=$tc timeout 120 python tools/time_resnet.py 512 12048 $cfg >> gpurun_out/r2f_time_resnet.log 2>&1; done; done
timeout 120 python tools/time_resnet.py 4096 12048 basic >> gpurun_out/r2f_time_resnet.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2f_resnet_launches.csv python tools/time_resnet.py 512 12048 basic > /dev/null 2>&1
cat gpurun_out/r2f_resnet_tests.log gpurun_out/r2f_dbg_basic.log gpurun_out/r2f_time_resnet.log
