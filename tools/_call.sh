set -x
T=r4m
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
grep -q failed gpurun_out/${T}_tests.log && exit 1
timeout -s KILL 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r4m_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_burst'], d['e2e']['value'], d['roofline_normalise']['frac'], d['clocks'])
P
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"fused01|conv_tc|conv_eo|conv_pair" -s 10 -c 10 -o gpurun_out/${T}_conv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
du -sh gpurun_out
