set -x
T=r4c
LE="timeout -s KILL 100 python tools/layer_events.py 4096 16000 3 12"
$LE warm > /dev/null 2>&1 || exit 1
for rep in 1 2 3 4; do
$LE oshift >> gpurun_out/${T}_layers.jsonl 2>/dev/null
RISER_E2_OSHIFT=0 $LE noshift >> gpurun_out/${T}_layers.jsonl 2>/dev/null
done
python - <<'P'
import json
for l in open('gpurun_out/r4c_layers.jsonl'):
    d=json.loads(l); print(d['tag'], {k:round(v,3) for k,v in d['layer_ms'].items() if int(k.split(':')[0])<=4}, round(d['conv_ms'],3))
P
