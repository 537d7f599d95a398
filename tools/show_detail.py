#!/usr/bin/env python
"""Print a bench.py --detail JSON: per-kernel totals then the top (kernel, shape) rows."""
import collections, json, sys
d = json.load(open(sys.argv[1]))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print("sum of kernel events: %.2f ms/step" % sum(x["ms_per_step"] for x in d))
agg = collections.Counter(); cnt = collections.Counter()
for x in d:
    k = x["name"].split("[")[0]; agg[k] += x["ms_per_step"]; cnt[k] += x["launches_per_step"]
for k, v in agg.most_common(n): print("%8.3f ms %6.0f x  %s" % (v, cnt[k], k))
print()
for x in d[:n]:
    print("%8.3f ms %5.1f x  %s  gf %.2f mb %.1f" % (x["ms_per_step"], x["launches_per_step"], x["name"], x["gflop_per_launch"], x["mbytes_per_launch"]))
