#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share."""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for x in csv.DictReader(lines):
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        rows.append((re.sub(r"\(.*", "", x["Kernel Name"]), v, x["Grid Size"], x["Block Size"]))
    return rows


def main():
    rows = load(sys.argv[1])
    tail = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    if tail:
        rows = rows[-tail:]
    agg = collections.OrderedDict()
    for n, v, g, b in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("launches %d  total %.1f us" % (len(rows), tot))
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-58s n=%4d total=%9.1f us avg=%8.2f us %5.1f%%" % (n[:58], a[0], a[1], a[1] / a[0], 100 * a[1] / tot))


if __name__ == "__main__":
    main()
