set -x
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_pipeline.py tests/test_gpu_resnet.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2l_tests.log
python bench.py --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 130 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
for lib in riser_b200/libriser_b200.so build_ab/replay.so; do echo $lib >> gpurun_out/r2l_time_resnet.log; RISER_B200_LIB=$lib timeout 120 python tools/time_resnet.py 512 12048 basic >> gpurun_out/r2l_time_resnet.log 2>&1; done
timeout 120 python tools/time_resnet.py 4096 12048 basic >> gpurun_out/r2l_time_resnet.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2l_resnet_launches.csv python tools/time_resnet.py 512 12048 basic > /dev/null 2>&1
cat gpurun_out/r2l_tests.log gpurun_out/r2l_time_resnet.log; cut -c1-160 gpurun_out/r2l_bench.json
