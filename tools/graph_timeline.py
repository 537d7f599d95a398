#!/usr/bin/env python
"""Timeline of ONE replay of the captured training step (CUPTI kernel records with start time and stream): per-stream busy time,
the gaps of the busiest stream (where the backbone waits for a lane or a lane's result), and the span of the step.

    python tools/graph_timeline.py [cfg2|cfg3] [min gap in us, default 8]
"""
import collections
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    min_gap = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0
    a = types.SimpleNamespace(unfused_optimizer=False, no_overlap=False, precision="parity", no_graph=False)
    import icl_b200
    icl_b200.set_precision("parity")
    dev = torch.device("cuda:0")
    tb = bench.TrainBench(a, wl, 0, 1, dev)
    tb.capture()
    for _ in range(3):
        tb.step(tb.x_dev, tb.y_dev)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        tb.step(tb.x_dev, tb.y_dev)
        torch.cuda.synchronize()
    path = os.path.join(ROOT, "gpurun_out", "timeline_%s.json" % wl)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    prof.export_chrome_trace(path)
    import json
    tr = json.load(open(path))
    evs = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    os.remove(path)
    evs.sort(key=lambda e: e["ts"])
    t0 = evs[0]["ts"]
    t1 = max(e["ts"] + e["dur"] for e in evs)
    print("# %s: one replay spans %.3f ms, %d kernels" % (wl, (t1 - t0) / 1e3, len(evs)))
    by = collections.defaultdict(list)
    for e in evs:
        by[e["args"].get("stream", e.get("tid"))].append(e)
    short = lambda n: n.replace("(anonymous namespace)::", "").split("(")[0].replace("void ", "")[:40]
    order = sorted(by, key=lambda s: -sum(e["dur"] for e in by[s]))
    for s in order:
        es = by[s]
        print("stream %-6s %4d kernels  busy %8.3f ms  first %8.3f  last end %8.3f   e.g. %s" % (
            s, len(es), sum(e["dur"] for e in es) / 1e3, (es[0]["ts"] - t0) / 1e3, (es[-1]["ts"] + es[-1]["dur"] - t0) / 1e3, short(es[0]["name"])))
    main_s = order[0]
    es = by[main_s]
    print("# gaps >= %.0f us on the busiest stream (%s):" % (min_gap, main_s))
    tot = 0.0
    for p, n in zip(es[:-1], es[1:]):
        gap = n["ts"] - (p["ts"] + p["dur"])
        if gap >= min_gap:
            tot += gap
            print("  at %8.3f ms  gap %7.1f us   after %-40s before %s" % ((p["ts"] + p["dur"] - t0) / 1e3, gap, short(p["name"]), short(n["name"])))
    print("# total of the listed gaps: %.3f ms" % (tot / 1e3))
    # coarse phases of the busiest stream: time stamps of a few marker kernels
    marks = ["conv3d_stem_fwd_k", "head1x1_fwd_k", "class_stats_row_fwd_k", "head1x1_bwd", "conv3d_stem_wgrad_k", "sgd_multi_k"]
    for m in marks:
        for e in evs:
            if m in e["name"]:
                print("# marker %-24s at %8.3f ms (stream %s)" % (m, (e["ts"] - t0) / 1e3, e["args"].get("stream", "?")))
                break
    tb.close()


if __name__ == "__main__":
    main()
