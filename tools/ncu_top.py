#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` output: top stalled SASS instructions per kernel."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for b in blocks[:1]:
    ix = {h: i for i, h in enumerate(b["hdr"])}
    data = b["data"]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    print(b["name"], "total samples", tot, "instructions", len(data))
    stalls = [h for h in b["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
    print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:n]:
        s = {h: int(r[ix[h]]) for h in stalls if int(r[ix[h]]) > 0}
        s = sorted(s.items(), key=lambda kv: -kv[1])[:3]
        print("%6s %9s  %-70s %s" % (r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:70], s))
