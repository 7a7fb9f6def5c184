"""Where do the single-thread roles of a conv kernel wait?  For one captured launch (ncu --set full --import-source on)
list the mbarrier try-wait spin loops (SYNCS.PHASECHK...TRYWAIT + the branch behind it) by barrier -- the offset in
ConvSmem names the barrier: a_full 0x00, a_empty 0x40, b_full 0x80, b_empty 0xc0, w_full 0x100, tmem_full 0x108,
tmem_full_ms 0x128 (conv_tc_kernel: per stage and sub-tile), tmem_empty 0x1a8.. -- with their stall samples, next to the kernel's headline counters.

    python tools/ncu_waits.py report.ncu-rep <launch index>
"""
import csv
import io
import re
import subprocess
import sys

rep, idx = sys.argv[1], sys.argv[2]
NAMES = [(0x1a8, "tmem_empty"), (0x128, "tmem_full_ms"), (0x108, "tmem_full"), (0x100, "w_full"), (0xc0, "b_empty"),
         (0x80, "b_full"), (0x40, "a_empty"), (0x00, "a_full")]


def barrier(off):
    for base, name in NAMES:
        if off >= base:
            return f"{name}[+0x{off - base:x}]"
    return hex(off)


out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-s", idx, "-c", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, ins, seen = None, [], set()
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) > 8 and r[0] not in seen:
        try:
            smp, n = int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        seen.add(r[0])
        ins.append((int(r[0], 16), r[1], smp, n))
total = sum(i[2] for i in ins)
print(f"launch {idx}: {len(ins)} SASS instructions, {total} stall samples (all warps)")
waits = {}
for k, (addr, text, smp, n) in enumerate(ins):
    m = re.search(r"TRYWAIT\s+P\d, \[R\d+\+URZ(?:\+0x([0-9a-f]+))?\]", text)
    if not m:
        continue
    off = int(m.group(1) or "0", 16)
    # the spin loop: this instruction + the (YIELD) / BRA around it
    loop = smp + sum(i[2] for i in ins[max(0, k - 1):k + 3] if ("BRA" in i[1] or "YIELD" in i[1]))
    w = waits.setdefault(barrier(off), [0, 0])
    w[0] += loop
    w[1] += n
for name, (smp, n) in sorted(waits.items(), key=lambda kv: -kv[1][0]):
    print(f"  wait on {name:22s} {smp:7d} samples ({100.0 * smp / max(total, 1):5.1f}% of all)   {n:10d} polls")
mma = sum(i[2] for i in ins if "UTC" in i[1] and "MMA" in i[1])
print(f"  samples on UTC*MMA instructions themselves: {mma}")
