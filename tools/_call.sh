set -x
T=r2o
LE="python tools/layer_events.py 4096 16000 3 12"
for rep in 1 2; do
RISER_PAIR=0 RISER_F8_FROM=6 $LE single_f6 >> gpurun_out/${T}_layers.jsonl 2>/dev/null
RISER_PAIR=1 RISER_F8_FROM=6 $LE pair_f6 >> gpurun_out/${T}_layers.jsonl 2>/dev/null
RISER_PAIR=1 RISER_F8_FROM=5 $LE pair_f5 >> gpurun_out/${T}_layers.jsonl 2>/dev/null
RISER_PAIR=0 RISER_F8_FROM=5 $LE single_f5 >> gpurun_out/${T}_layers.jsonl 2>/dev/null
done
RISER_PAIR=1 RISER_F8_FROM=6 RISER_PAIR_NTILE=192 $LE pair_f6_nt192 >> gpurun_out/${T}_layers.jsonl 2>/dev/null
for v in 1 2 16 32; do RISER_B200_LIB=build_ab/dbg$v.so RISER_PAIR=0 RISER_F8_FROM=6 $LE dbg$v >> gpurun_out/${T}_layers.jsonl 2>/dev/null; done
for rep in 1 2; do
RISER_PAIR=0 RISER_F8_FROM=6 python bench.py --no-cpu-baseline --steps 20 >> gpurun_out/${T}_bench_single_f6.jsonl 2>/dev/null
RISER_PAIR=1 RISER_F8_FROM=6 python bench.py --no-cpu-baseline --steps 20 >> gpurun_out/${T}_bench_pair_f6.jsonl 2>/dev/null
done
cat gpurun_out/${T}_layers.jsonl | cut -c1-400
cut -c1-160 gpurun_out/${T}_bench_*.jsonl
