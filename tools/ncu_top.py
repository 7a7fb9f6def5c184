"""Top stall locations of one captured launch (ncu --set full --import-source on): SASS instructions ranked by stall
samples with their dominant stall reason, plus totals per reason.  usage: python tools/ncu_top.py rep.ncu-rep <launch> [n=40]"""
import csv
import io
import subprocess
import sys

rep, idx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-s", idx, "-c", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
ins = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        ins.append(r)
si = hdr.index("# Samples")
reasons = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si]) for r in ins)
per = {hdr[i]: sum(int(r[i] or 0) for r in ins) for i in reasons}
print(f"{len(ins)} instructions, {tot} samples; by reason:", ", ".join(f"{k[6:]} {v}" for k, v in sorted(per.items(), key=lambda kv: -kv[1]) if v))
order = sorted(range(len(ins)), key=lambda k: -int(ins[k][si]))[:n]
for k in sorted(order):
    r = ins[k]
    top = max(reasons, key=lambda i: int(r[i] or 0))
    print(f"{k:5d} {int(r[si]):6d} {100.0 * int(r[si]) / max(tot, 1):5.1f}%  exec {r[hdr.index('Instructions Executed')]:>9s}  {hdr[top][6:]:12s} {r[1].strip()[:100]}")
