"""Per-source-line and per-role breakdown of one ncu capture (--import-source on):
instructions executed, stall samples and stall reasons, grouped by line ranges.
usage: python tools/ncu_roles.py report.ncu-rep file.cu name:lo-hi [name:lo-hi ...]"""
import csv
import io
import subprocess
import sys

rep, src = sys.argv[1], sys.argv[2]
roles = []
for spec in sys.argv[3:]:
    name, rng = spec.split(":")
    lo, hi = rng.split("-")
    roles.append((name, int(lo), int(hi)))
import os
sel = (["-s", os.environ["NCU_LAUNCH"], "-c", "1"] if os.environ.get("NCU_LAUNCH") else [])   # one launch of the report
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + sel,
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, hdr = None, None
agg, stall, lines = {}, {}, {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 8 and r[0].isdigit() and r[2] == "-":
        ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
        try:
            n, smp = int(r[ie]), int(r[ss])
        except ValueError:
            continue
        ln = int(r[0])
        role = "other:" + cur
        if cur == src:
            role = "other"
            for name, lo, hi in roles:
                if lo <= ln <= hi:
                    role = name
        a = agg.setdefault(role, [0, 0])
        a[0] += n
        a[1] += smp
        lines[(cur, ln)] = (n, smp, r[1].strip()[:90])
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h:
                try:
                    stall[(role, h)] = stall.get((role, h), 0) + int(r[i])
                except ValueError:
                    pass
ti = sum(v[0] for v in agg.values())
ts = sum(v[1] for v in agg.values())
print(f"total warp instructions {ti}, samples {ts}")
for role, (n, smp) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    t = sum(v for (rr, h), v in stall.items() if rr == role) or 1
    top = sorted([(h[6:], round(v / t * 100)) for (rr, h), v in stall.items() if rr == role and v / t > 0.04],
                 key=lambda x: -x[1])
    print(f"{role:28s} inst {n / ti * 100:5.1f}%  samples {smp / ts * 100:5.1f}%  {top}")
print("top lines by samples:")
for k, v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"  {k[0]}:{k[1]:<5d} inst {v[0] / ti * 100:5.1f}%  smp {v[1] / ts * 100:5.1f}%  {v[2]}")
