set -x
T=r2w
timeout -s KILL 300 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "every_layer" 2>&1 | tail -5 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
LE="timeout -s KILL 200 python tools/layer_events.py 4096 16000 3 12"
$LE warm > /dev/null 2>&1
for rep in 1 2; do
$LE res5 >> gpurun_out/${T}_layers.jsonl 2>/dev/null
RISER_PAIR_RESIDENT=0 $LE nores >> gpurun_out/${T}_layers.jsonl 2>/dev/null
done
timeout -s KILL 300 python bench.py --no-cpu-baseline --steps 20 >> gpurun_out/${T}_bench.jsonl 2>/dev/null
