set -x
T=r3e
LE="timeout -s KILL 200 python tools/layer_events.py 4096 16000 3 12"
$LE warm > /dev/null 2>&1
for d in 0 1 2 3 0; do
RISER_PAIR_DBG_SKIP=$d $LE skip$d >> gpurun_out/${T}_layers.jsonl 2>/dev/null
done
python - <<'P'
import json
for l in open('gpurun_out/r3e_layers.jsonl'):
    d=json.loads(l); print(d['tag'], {k:round(v,3) for k,v in d['layer_ms'].items() if int(k.split(':')[0])>=5}, round(d['conv_ms'],3))
P
