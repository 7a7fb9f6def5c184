"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries kept under
profiles/ (the .ncu-rep files themselves are scratch).

  python tools/summarise_ncu.py launches <launches.csv>          # per-kernel time shares
  python tools/summarise_ncu.py raw <file.ncu-rep>               # key metrics per captured launch
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def short(name):
    name = name.replace("riser::<unnamed>::", "")
    return name.split("(")[0][:60]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        k = short(r[ki])
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + float(r[vi].replace(",", "")))
    total = sum(t for _, t in agg.values())
    print(f"# {path}: {sum(n for n, _ in agg.values())} launches, {total / 1e6:.3f} ms total (ncu, serialised, cold cache)")
    print(f"{'kernel':62s} {'launches':>8s} {'total us':>10s} {'share':>7s} {'avg us':>9s}")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:62s} {n:8d} {t / 1e3:10.1f} {100 * t / total:6.1f}% {t / n / 1e3:9.1f}")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    cols = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    print(f"# {path}: {len(rows) - 2} captured launches")
    for n, r in enumerate(rows[2:]):
        print(f"[{n}] {short(r[ki])}")
        for k, i in cols:
            print(f"    {k:66s} {r[i]:>14s} {units[i]}")
        try:
            rd, wr = float(r[hdr.index('dram__bytes_read.sum')]), float(r[hdr.index('dram__bytes_write.sum')])
            print(f"    {'dram read+write (traffic)':66s} {rd + wr:14.3f} {units[hdr.index('dram__bytes_read.sum')]}")
        except Exception:
            pass


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
