set -x
T=r3q
timeout -s KILL 300 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
grep -q failed gpurun_out/${T}_tests.log && exit 1
for s in 4 1 4 1; do
RISER_UPLOAD_SLICES=$s timeout -s KILL 150 python tools/live_breakdown.py 3000 1 2>/dev/null | tail -1 | sed "s/^/slices=$s /" >> gpurun_out/${T}_lb.log
done
for t in 4 8; do
RISER_PACK_THREADS=$t timeout -s KILL 150 python tools/live_breakdown.py 3000 1 2>/dev/null | tail -1 | sed "s/^/threads=$t /" >> gpurun_out/${T}_lb.log
done
RISER_UPLOAD_SLICES=1 timeout -s KILL 150 python tools/live_breakdown.py 512 2 2>/dev/null | tail -1 | sed "s/^/slices=1 /" >> gpurun_out/${T}_lb.log
timeout -s KILL 150 python tools/live_breakdown.py 512 2 2>/dev/null | tail -1 | sed "s/^/slices=4 /" >> gpurun_out/${T}_lb.log
cut -c1-330 gpurun_out/${T}_lb.log
