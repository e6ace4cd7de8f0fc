#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list per kernel: launches, total time, share.
usage: ncu_launch_list.py launches.csv [steps]   (steps = bench steps the run made, to print per-step figures)"""
import csv, re, sys
from collections import OrderedDict
path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lines = [l for l in open(path, errors="replace") if l.startswith('"')]
rows = list(csv.reader(lines))
hdr = rows[0]
ik, iv, iu, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}
agg = OrderedDict()
total = 0.0
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    us = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
    name = re.sub(r"\(.*", "", r[ik])
    name = re.sub(r"^void ", "", name).replace("sm100::", "").replace("<unnamed>::", "")
    e = agg.setdefault(name, [0, 0.0])
    e[0] += 1
    e[1] += us
    total += us
print("# %s: %d launches, %.3f ms under ncu (serialised, cold caches); %d bench steps in the run" % (path, sum(e[0] for e in agg.values()), total * 1e-3, steps))
print("%-100s %8s %12s %7s" % ("kernel", "launches", "total_ms", "share"))
for k, e in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-100s %8d %12.3f %6.1f%%" % (k[:100], e[0], e[1] * 1e-3, 100.0 * e[1] / total))
umma = sum(e[1] for k, e in agg.items() if "umma_" in k)
print("# tcgen05 contraction kernels (umma_*): %.3f ms = %.1f%% of the launch time" % (umma * 1e-3, 100.0 * umma / total))
