#!/usr/bin/env python
"""Per-kernel totals and shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/launch_shares.py profiles/r02_launches_bench.csv [more.csv ...]"""
import collections
import csv
import re
import sys

UNIT_US = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", "")) * UNIT_US.get(r[iu], 1.0)
        except ValueError:
            continue
        a = agg.setdefault(re.sub(r"\(.*", "", r[ik])[:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    print("== %s: %d launches, %.1f us of kernel time (cold-cache, serialised: compare shares)" % (path, sum(a[0] for a in agg.values()), total))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:16]:
        print("  %-72s n=%5d  %11.1f us  %8.1f us/launch  %5.1f %%" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / total))
