set -x
T=r2r
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${T}_tests.log
LE="python tools/layer_events.py 4096 16000 3 12"
for rep in 1 2; do
$LE pair_ms2_dual >> gpurun_out/${T}_layers.jsonl 2>/dev/null
RISER_PAIR_MS=1 $LE pair_ms1 >> gpurun_out/${T}_layers.jsonl 2>/dev/null
done
RISER_PAIR_FROM=5 $LE pair_from5 >> gpurun_out/${T}_layers.jsonl 2>/dev/null
RISER_PAIR_FROM=5 RISER_PAIR_MS=1 $LE pair_from5_ms1 >> gpurun_out/${T}_layers.jsonl 2>/dev/null
RISER_PAIR_NTILE=192 $LE pair_nt192 >> gpurun_out/${T}_layers.jsonl 2>/dev/null
RISER_F8_FROM=6 $LE f8from6 >> gpurun_out/${T}_layers.jsonl 2>/dev/null
for rep in 1 2; do
python bench.py --no-cpu-baseline --steps 20 >> gpurun_out/${T}_bench.jsonl 2>/dev/null
RISER_PAIR_MS=1 python bench.py --no-cpu-baseline --steps 20 >> gpurun_out/${T}_bench_ms1.jsonl 2>/dev/null
done
ncu --set full --clock-control none --import-source on -k regex:"conv_eo" -c 3 -o gpurun_out/${T}_eo python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
cat gpurun_out/${T}_tests.log
