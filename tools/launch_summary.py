#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel."""
import collections
import csv
import sys

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0  # launches to drop from the front (warm-up, probes)
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
hdr = rows[0]
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = []
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
    launches.append((r[ik].split("(")[0], v * scale))
launches = launches[skip:]
tot = sum(t for _, t in launches)
agg = collections.OrderedDict()
for k, t in launches:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += t
print("%d launches, %.3f ms of kernel time (cold-cache, serialised under ncu: compare shares)" % (len(launches), tot / 1e3))
print("%-52s %8s %12s %12s %8s" % ("kernel", "launches", "total us", "mean us", "share"))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-52s %8d %12.1f %12.1f %7.2f%%" % (k[:52], n, t, t / n, 100.0 * t / tot))
