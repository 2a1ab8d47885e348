#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except (ValueError, IndexError):
        continue
    u = r[ui]
    ms = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v if u.startswith("m") else v * 1e3
    name = r[ki].split("(")[0].replace("void ", "").replace("tpc::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
total = sum(a[1] for a in agg.values())
for name, (c, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{ms:10.3f} ms {100 * ms / total:5.1f}% {c:6d} x  {name[:90]}")
print(f"{total:10.3f} ms total")
