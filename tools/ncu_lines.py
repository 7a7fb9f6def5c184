"""Per-source-line totals of one captured launch (ncu --set full --import-source on, built with -lineinfo): warp
instructions executed, stall samples (with the dominant reason) and shared-memory wavefronts, for the lines that carry
the most of either.  usage: python tools/ncu_lines.py rep.ncu-rep <launch> [n=40] [file-substring]"""
import csv
import io
import subprocess
import sys

rep, idx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
only = sys.argv[4] if len(sys.argv) > 4 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-s", idx, "-c", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, fname = None, ""
lines = {}
for r in rows:
    if r and r[0] in ("File Name", "File Path"):
        fname = r[1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if not hdr or len(r) < len(hdr) or not r[0].isdigit() or (only and only not in fname):
        continue
    try:
        r = [r[0], ",".join(r[1:len(r) - len(hdr) + 2])] + r[len(r) - len(hdr) + 2:]   # (commas inside the source text)
        ex = int(r[hdr.index("Instructions Executed")] or 0)
        sm = int(r[hdr.index("# Samples")] or 0)
        wf = int(r[hdr.index("L1 Wavefronts Shared")] or 0)
    except ValueError:
        continue
    key = (fname.split("/")[-1], int(r[0]))
    d = lines.setdefault(key, {"src": r[1].strip(), "ex": 0, "sm": 0, "wf": 0, "why": {}})
    d["ex"] += ex
    d["sm"] += sm
    d["wf"] += wf
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h and r[i]:
            d["why"][h[6:]] = d["why"].get(h[6:], 0) + int(r[i])
tex, tsm, twf = (sum(d[k] for d in lines.values()) for k in ("ex", "sm", "wf"))
print(f"launch {idx}: {tex} warp instructions, {tsm} stall samples, {twf} shared wavefronts over {len(lines)} source lines")
top = set(sorted(lines, key=lambda k: -lines[k]["ex"])[:n]) | set(sorted(lines, key=lambda k: -lines[k]["sm"])[:n])
for key in sorted(top):
    d = lines[key]
    why = max(d["why"], key=d["why"].get) if d["why"] else ""
    print(f"{key[0]}:{key[1]:5d} instr {100.0 * d['ex'] / max(tex, 1):5.1f}%  samples {100.0 * d['sm'] / max(tsm, 1):5.1f}% "
          f"({why:14s}) wavefronts {100.0 * d['wf'] / max(twf, 1):5.1f}%  {d['src'][:90]}")
