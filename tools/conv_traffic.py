"""Regenerate profiles/conv_stack_traffic.json (what bench.py reports as roofline.traffic) from an `ncu --set full`
capture of the conv launches of ONE forward:

    ncu --set full --clock-control none --import-source on -k regex:"fused01|conv_tc|conv_eo|conv_pair" -s 11 -c 11 \
        -o gpurun_out/<name> python bench.py --steps 1 --warmup 1 --no-cpu-baseline          (under gpurun)
    python tools/conv_traffic.py gpurun_out/<name>.ncu-rep B4096_L16000_p3

Sum of dram__bytes_read.sum + dram__bytes_write.sum over the captured launches, with per-launch detail, the commit
the library was built from and the date, so that the figure cannot go stale silently.
"""
import csv
import datetime
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, key = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
ti = hdr.index("gpu__time_duration.sum")
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
per, total = [], 0.0
for r in rows[2:]:
    b = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
    total += b
    per.append({"kernel": r[ki].split("<unnamed>::")[-1].split("(")[0], "ms": float(r[ti]) * (1.0 if units[ti] == "ms" else 1e-3),
                "dram_bytes": b})
commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
dirty = bool(subprocess.run(["git", "-C", ROOT, "status", "--porcelain", "riser_b200/csrc"], capture_output=True,
                            text=True).stdout.strip())
path = os.path.join(ROOT, "profiles", "conv_stack_traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[key] = total
data["_source"] = (f"{os.path.basename(rep)}: {len(per)} launches of one forward, ncu --set full, "
                   f"commit {commit}{'+uncommitted csrc changes' if dirty else ''}, "
                   f"{datetime.date.today().isoformat()} (tools/conv_traffic.py)")
data.setdefault("_detail", {})[key] = per
data["_note"] = ("sum of dram__bytes_read.sum + dram__bytes_write.sum over the conv launches of one forward "
                 "(layer 0 fused into layer 1's launch); keys B<batch>_L<samples>_p<precision mode>")
json.dump(data, open(path, "w"), indent=1)
print(f"{key}: {total / 1e9:.3f} GB over {len(per)} launches -> {path}")
