set -x
T=r4d
timeout -s KILL 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r4d_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_burst'], d['e2e']['value'], d['roofline_normalise']['frac'], d['clocks'])
P
timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:normalise -s 3 -c 1 -o gpurun_out/${T}_norm python tools/time_normalise.py > /dev/null 2>&1
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout -s KILL 600 python tools/parity_sweep.py 8192 91 > gpurun_out/${T}_parity_sweep.json 2> gpurun_out/${T}_parity.err
cut -c1-900 gpurun_out/${T}_parity_sweep.json
