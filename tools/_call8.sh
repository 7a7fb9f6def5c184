set -x
T=r3c
nvidia-smi -L | head -8 > gpurun_out/${T}_gpus.txt
timeout -s KILL 900 python -m pytest tests/test_gpu_shard.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${T}_shard_tests.log
cat gpurun_out/${T}_shard_tests.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -s KILL 600 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/${T}_bench_8gpu.json 2> gpurun_out/${T}_8.err
tail -c 300 gpurun_out/${T}_bench_8gpu.json
timeout -s KILL 600 $TR --nproc-per-node 8 --master-port 29532 bench.py --gpus 8 --reads 1000000 > gpurun_out/${T}_reads1M_8gpu.json 2>> gpurun_out/${T}_8.err
cat gpurun_out/${T}_reads1M_8gpu.json | cut -c1-1500
timeout -s KILL 600 python bench.py --gpus 1 --reads 1000000 > gpurun_out/${T}_reads1M_1gpu.json 2>> gpurun_out/${T}_8.err
cat gpurun_out/${T}_reads1M_1gpu.json | cut -c1-800
tail -5 gpurun_out/${T}_8.err
