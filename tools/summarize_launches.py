"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share of the captured window).
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt"""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(r[ui], 1.0)
        a = agg.setdefault(name, [0, 0.0, r[bi]])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows) - 1} launches, {tot / 1e3:.1f} us total (cold-cache, serialised: compare SHARES, not absolutes)")
    print(f"{'kernel':58s} {'launches':>8s} {'total_us':>10s} {'avg_us':>8s} {'share':>6s}  block")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:58s} {v[0]:8d} {v[1] / 1e3:10.1f} {v[1] / 1e3 / v[0]:8.1f} {v[1] / tot:6.3f}  {v[2]}")


if __name__ == "__main__":
    main(sys.argv[1])
