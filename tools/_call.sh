set -x
T=r2u
timeout -s KILL 240 python -m pytest "tests/test_gpu_model.py::test_every_layer_against_oracle[2-3]" -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${T}_tests_b.log
cat gpurun_out/${T}_tests_b.log
if grep -q passed gpurun_out/${T}_tests_b.log; then
timeout -s KILL 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${T}_tests.log
LE="timeout -s KILL 200 python tools/layer_events.py 4096 16000 3 12"
for rep in 1 2; do
$LE kc >> gpurun_out/${T}_layers.jsonl 2>/dev/null
RISER_KC=0 $LE nokc >> gpurun_out/${T}_layers.jsonl 2>/dev/null
done
timeout -s KILL 300 python bench.py --no-cpu-baseline --steps 20 >> gpurun_out/${T}_bench.jsonl 2>/dev/null
RISER_KC=0 timeout -s KILL 300 python bench.py --no-cpu-baseline --steps 20 >> gpurun_out/${T}_bench_nokc.jsonl 2>/dev/null
timeout -s KILL 300 python bench.py --no-cpu-baseline --steps 20 >> gpurun_out/${T}_bench.jsonl 2>/dev/null
fi
cat gpurun_out/${T}_tests.log
