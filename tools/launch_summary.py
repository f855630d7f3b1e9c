"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file x.csv`) per kernel.

  python tools/launch_summary.py gpurun_out/launches.csv "command that was profiled" > profiles/x_summary.txt
"""
import collections
import csv
import sys


def main():
    path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if not l.startswith("==")) if r]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
    agg = collections.OrderedDict()
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) <= max(ki, vi, ui):
            continue
        ms = float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-6)
        a = agg.setdefault(r[ki], [0, 0.0, 0.0])
        a[0] += 1; a[1] += ms; a[2] = max(a[2], ms)
    total = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none : {cmd}")
    print(f"# {n} launches, {total:.1f} ms of kernel time (cold-cache, serialised: shares, not absolutes)")
    for k, (c, t, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:100]:100s} n={c:5d} total {t:10.2f} ms  avg {1e3 * t / c:10.1f} us  max {1e3 * mx:10.1f} us  {100 * t / total:5.1f}%")


if __name__ == "__main__":
    main()
