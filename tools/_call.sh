set -x
T=r3b
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
timeout -s KILL 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 600 gpurun_out/${T}_bench.json
timeout -s KILL 600 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_bench_ref.json | cut -c1-400
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"fused01|conv_tc|conv_eo|conv_pair" -s 10 -c 10 -o gpurun_out/${T}_conv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/ | tail -8
