#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.

    python tools/ncu_launch_summary.py gpurun_out/launches.csv [skip_launches] > profiles/rNN_launches_summary.txt
"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1], newline="") as f:
    rd = csv.reader(f)
    hdr = None
    for r in rd:
        if hdr is None:
            if r and r[0] == "ID":
                hdr = {h: i for i, h in enumerate(r)}
            continue
        if len(r) >= len(hdr) and r[hdr["Metric Name"]] == "gpu__time_duration.sum":
            v = float(r[hdr["Metric Value"]].replace(",", ""))
            unit = r[hdr["Metric Unit"]]
            us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
            rows.append((int(r[hdr["ID"]]), r[hdr["Kernel Name"]], us))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = [r for r in rows if r[0] >= skip]
agg = collections.OrderedDict()
for _, name, us in rows:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("# %d launches, %.3f ms total device time (cold-cache, serialised under ncu: compare SHARES)" % (len(rows), tot / 1e3))
print("%-70s %8s %12s %7s" % ("kernel", "launches", "total_us", "share"))
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s %8d %12.1f %6.1f%%" % (k[:70], n, us, 100 * us / tot))
