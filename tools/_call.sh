set -x
T=r3s
timeout -s KILL 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r3s_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_burst'], d['e2e']['value'], d['roofline_normalise']['frac'], d['clocks'])
P
timeout -s KILL 600 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:normalise -s 3 -c 1 -o gpurun_out/${T}_norm python tools/time_normalise.py > /dev/null 2>&1
timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:polya -s 2 -c 1 -o gpurun_out/${T}_polya python tools/time_preprocess.py > /dev/null 2>&1
timeout -s KILL 300 ncu --set full --clock-control none -k regex:"res_tc|stem_pool" -s 13 -c 13 -o /tmp/${T}_resnet python tools/time_resnet.py 512 12048 basic > /dev/null 2>&1
python tools/summarise_ncu.py raw /tmp/${T}_resnet.ncu-rep > gpurun_out/${T}_resnet_full.txt 2>&1
timeout -s KILL 100 python tools/time_preprocess.py 2>/dev/null | tail -2 > gpurun_out/${T}_pre.log
timeout -s KILL 100 python tools/time_normalise.py 2>/dev/null | tail -1 >> gpurun_out/${T}_pre.log
cat gpurun_out/${T}_pre.log
du -sh gpurun_out
