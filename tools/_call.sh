set -x
T=r3j
timeout -s KILL 120 python tools/batch_invariance.py 96 > gpurun_out/${T}_inv.log 2>&1
tail -25 gpurun_out/${T}_inv.log
timeout -s KILL 120 python tools/batch_invariance.py 96 RISER_FUSE23=0 2>&1 | tail -4 > gpurun_out/${T}_inv_nofuse23.log
cat gpurun_out/${T}_inv_nofuse23.log
timeout -s KILL 120 python tools/batch_invariance.py 96 RISER_PAIR=0 2>&1 | tail -4 > gpurun_out/${T}_inv_nopair.log
cat gpurun_out/${T}_inv_nopair.log
timeout -s KILL 120 python tools/batch_invariance.py 96 RISER_KC=0 2>&1 | tail -4 > gpurun_out/${T}_inv_nokc.log
cat gpurun_out/${T}_inv_nokc.log
timeout -s KILL 120 python tools/batch_invariance.py 96 RISER_LIVE_GRAPHS=0 2>&1 | tail -4 > gpurun_out/${T}_inv_nograph.log
cat gpurun_out/${T}_inv_nograph.log
