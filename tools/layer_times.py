"""Per-layer kernel durations of one forward from an ncu launch list (csv written with
--metrics gpu__time_duration.sum ... --csv --log-file X).  usage: python tools/layer_times.py X [skip]"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 11
by = {}
for r in rows:
    by.setdefault(int(r[0]), {"name": r[4]})[r[-3]] = float(r[-1].replace(",", ""))
ids = sorted(by)[skip:skip + 11]
t = [by[i]["gpu__time_duration.sum"] / 1e6 for i in ids]
print(" ".join(f"{x:.3f}" for x in t), "| sum %.3f ms" % sum(t))
