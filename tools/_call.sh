set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2a_tests.log
python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_st256.json 2> gpurun_out/r2a_bench_st256.err
RISER_B200_LIB=build_ab/st128.so python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_st128.json 2> gpurun_out/r2a_bench_st128.err
python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_st256b.json 2>> gpurun_out/r2a_bench_st256.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
RISER_B200_LIB=build_ab/st128.so ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches_st128.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fused01|conv_tc|conv_eo" -s 11 -c 11 -o gpurun_out/r2a_conv_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/r2a_ncu.err
tail -3 gpurun_out/r2a_tests.log; cat gpurun_out/r2a_bench_st256.json gpurun_out/r2a_bench_st128.json gpurun_out/r2a_bench_st256b.json | cut -c1-400
