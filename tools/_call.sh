set -x
T=r3d
timeout -s KILL 600 python -m pytest tests/test_gpu_convnet_generic.py tests/test_gpu_resnet.py tests/test_gpu_dropin.py -m gpu -x -q -rP 2>&1 | grep -v "^$" | tail -25 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
timeout -s KILL 300 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -s 2>&1 | grep -E "forgiven|passed|failed" > gpurun_out/${T}_pipeline.log
cat gpurun_out/${T}_pipeline.log
timeout -s KILL 600 python bench.py --gpus 1 --reads 1000000 > gpurun_out/${T}_reads1M_1gpu.json 2> gpurun_out/${T}_reads.err
cut -c1-300 gpurun_out/${T}_reads1M_1gpu.json; tail -3 gpurun_out/${T}_reads.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r3d_reads1M_1gpu.json'))
print(d['value'], d['ms_per_step'], d['sharding'])
P
