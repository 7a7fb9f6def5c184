"""SASS census of libriser_b200.so: per kernel, how many tcgen05 / TMEM / TMA instructions the binary holds
(B200_PROFILING.md "What proves a Blackwell-native kernel").  Runs without a GPU.

    python tools/sass_census.py [lib.so] > profiles/r2_sass_census.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "riser_b200", "libriser_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCHMMA.2CTA", "UTCQMMA.2CTA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG",
             "UBLKCP", "UTCCP", "STG.E.ENL2.256", "HMMA", "SYNCS"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True,
                          text=True).stdout.splitlines()
names = iter(demangle)
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(names, m.group(1))
        cur = re.sub(r"\((int|bool)\)", "", cur)
        cur = re.sub(r"\(.*", "", cur).replace("riser::(anonymous namespace)::", "").replace("riser::<unnamed>::", "").replace("void ", "")
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    per[cur]["_all"] += 1
    for k in MNEMONICS:
        if op == k or op.startswith(k + ".") or (k.endswith(".256") and op.startswith(k)):
            per[cur][k] += 1
    if ".2CTA" in op:
        per[cur][op.split(".")[0] + ".2CTA"] += 1
tot = collections.Counter()
cols = [k for k in MNEMONICS if any(c[k] for c in per.values())]
print(f"# SASS census of {os.path.relpath(lib, ROOT)} (cuobjdump -sass; {len(per)} kernels)")
print(f"{'kernel':58s} {'instr':>7s} " + " ".join(f"{k[:12]:>12s}" for k in cols))
for name, c in per.items():
    if not any(c[k] for k in cols if k != "SYNCS"):
        continue
    print(f"{name[:58]:58s} {c['_all']:7d} " + " ".join(f"{c[k]:12d}" for k in cols))
    tot.update(c)
print(f"{'total (all kernels with any of the above)':58s} {tot['_all']:7d} " + " ".join(f"{tot[k]:12d}" for k in cols))
