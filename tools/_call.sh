set -x
T=r4j
for v in cvt6 cvt9; do
RISER_B200_LIB=build_ab/$v.so timeout -s KILL 200 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "every_layer" 2>&1 | tail -2 | sed "s/^/$v /" >> gpurun_out/${T}_tests.log
done
cat gpurun_out/${T}_tests.log
grep -q failed gpurun_out/${T}_tests.log && exit 1
LE="timeout -s KILL 100 python tools/layer_events.py 4096 16000 3 12"
$LE warm > /dev/null 2>&1 || exit 1
for rep in 1 2 3; do
$LE cvt4 >> gpurun_out/${T}_layers.jsonl 2>/dev/null
RISER_B200_LIB=build_ab/cvt6.so $LE cvt6 >> gpurun_out/${T}_layers.jsonl 2>/dev/null
RISER_B200_LIB=build_ab/cvt9.so $LE cvt9 >> gpurun_out/${T}_layers.jsonl 2>/dev/null
done
python - <<'P'
import json
for l in open('gpurun_out/r4j_layers.jsonl'):
    d=json.loads(l); print(d['tag'], {k:round(v,3) for k,v in d['layer_ms'].items() if int(k.split(':')[0]) <=2}, round(d['conv_ms'],3))
P
