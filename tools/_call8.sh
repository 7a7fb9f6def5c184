set -x
T=r3t
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -s KILL 300 $TR --nproc-per-node 8 --master-port 29532 bench.py --gpus 8 --reads 1000000 > gpurun_out/${T}_reads1M_8gpu.json 2> gpurun_out/${T}_8.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r3t_reads1M_8gpu.json')); print(d['value'], d['ms_per_step'], d['sharding'])
P
timeout -s KILL 300 $TR --nproc-per-node 2 --master-port 29533 bench.py --gpus 2 --reads 1000000 > gpurun_out/${T}_reads1M_2gpu.json 2>> gpurun_out/${T}_8.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r3t_reads1M_2gpu.json')); print(d['value'], d['ms_per_step'], d['sharding']['host_gb_per_s_per_rank'])
P
nproc; tail -3 gpurun_out/${T}_8.err
