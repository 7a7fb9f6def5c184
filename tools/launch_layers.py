"""Per-launch durations of the LAST complete forward in an ncu launch list (csv from
--metrics gpu__time_duration.sum --csv --log-file X): one line per kernel of the conv stack, in launch order.
usage: python tools/launch_layers.py X [X ...]"""
import csv
import sys

for path in sys.argv[1:]:
    rows = []
    for r in csv.reader(open(path, errors="replace")):
        if len(r) > 10 and r[0].isdigit() and r[-3] == "gpu__time_duration.sum":
            rows.append((r[4].split("(")[0].replace("void riser::<unnamed>::", ""), float(r[-1].replace(",", ""))))
    unit = 1e6   # ns -> ms
    # a forward = the launches between two fused01 launches
    starts = [i for i, (n, _) in enumerate(rows) if n.startswith("fused01")]
    if len(starts) < 2:
        print(path, "no complete forward")
        continue
    a, b = starts[-2], starts[-1]
    fw = [(n, t) for n, t in rows[a:b] if n.startswith(("fused01", "conv_"))]
    print(f"# {path}")
    print("  " + " ".join(f"{t / unit:.3f}" for _, t in fw) + f" | sum {sum(t for _, t in fw) / unit:.3f} ms")
    print("  " + " ".join(n.split("<")[0].replace("_kernel", "").replace("conv_", "") for n, _ in fw))
