#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.
usage: python profiles/summarize.py profiles/launches_rNN.csv > table.md"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            name = r[hdr.index("Kernel Name")]
            v = float(r[hdr.index("Metric Value")])
            u = r[hdr.index("Metric Unit")]
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
            name = name.replace("<unnamed>::", "").split("(")[0]
            a = agg.setdefault(name, [0, 0.0, 0.0])
            a[0] += 1
            a[1] += v
            a[2] = max(a[2], v)
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | max ms | share |")
    print("|---|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.3f | %.3f | %.3f |" % (k[:70], a[0], a[1], a[2], a[1] / tot))
    print("\n%d launches, %.3f ms in total" % (sum(a[0] for a in agg.values()), tot))


if __name__ == "__main__":
    main(sys.argv[1])
