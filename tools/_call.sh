set -x
T=r4n
RISER_F01_QUAD=1 timeout -s KILL 200 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "every_layer" 2>&1 | tail -2 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
grep -q passed gpurun_out/${T}_tests.log || exit 1
grep -q failed gpurun_out/${T}_tests.log && exit 1
LE="timeout -s KILL 100 python tools/layer_events.py 4096 16000 3 12"
$LE warm > /dev/null 2>&1 || exit 1
for rep in 1 2 3 4; do
RISER_F01_QUAD=1 $LE quad >> gpurun_out/${T}_layers.jsonl 2>/dev/null
$LE noquad >> gpurun_out/${T}_layers.jsonl 2>/dev/null
done
python - <<'P'
import json
for l in open('gpurun_out/r4n_layers.jsonl'):
    d=json.loads(l); print(d['tag'], {k:round(v,3) for k,v in d['layer_ms'].items() if int(k.split(':')[0]) <=2}, round(d['conv_ms'],3))
P
