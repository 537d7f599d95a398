#!/usr/bin/env python
"""Micro-benchmark of the weight-streaming skinny GEMMs of mlp2 (rows x 13824 x 13824 and the smaller levels) and of the fused
rank-R SGD update; prints achieved GB/s of the fp32 weight stream against the measured HBM peak.  Used under ncu."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icl_b200 import ops  # noqa: E402
from icl_b200.ops import P, c_f, c_int, call  # noqa: E402


def timeit(fn, iters=6):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk)).get("hbm_gbs", 6546.2) if os.path.exists(pk) else 6546.2
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    g = torch.Generator(device="cuda").manual_seed(1)
    for rows, n in ((16, 13824), (128, 13824), (32, 1728), (64, 216)):
        x = torch.randn(rows, n, device="cuda", generator=g)
        w = torch.randn(n, n, device="cuda", generator=g) * 0.01
        b = torch.zeros(n, device="cuda")
        mb = n * n * 4 / 1e6
        if only in ("", "fwd"):
            ms = timeit(lambda: ops.linear_fwd(x, w, b, 1, want_pre=True))
            print("linear_fwd  %4dx%dx%d  %8.3f ms  %7.0f GB/s (%4.1f%% of %.0f)" % (rows, n, n, ms, mb / ms, 100 * mb / ms / peak, peak), flush=True)
        if only in ("", "dgrad"):
            ms = timeit(lambda: ops.linear_dgrad(x, w))
            print("linear_dgrad %4dx%dx%d %8.3f ms  %7.0f GB/s (%4.1f%%)" % (rows, n, n, ms, mb / ms, 100 * mb / ms / peak), flush=True)
        if only in ("", "sgd") and rows == 16:
            m = torch.zeros_like(w)
            lr = torch.full((1,), 0.01, device="cuda")
            dy = torch.randn(rows, n, device="cuda", generator=g)
            ms = timeit(lambda: call("icl_sgd_factored", c_int(rows), c_int(n), c_int(n), P(dy), P(x), P(w), P(m), P(lr), c_f(0.9), c_f(1e-4)))
            print("sgd_factored R%d %dx%d      %8.3f ms  %7.0f GB/s of 16 B/param (%4.1f%%)" % (rows, n, n, ms, 4 * mb / ms, 100 * 4 * mb / ms / peak), flush=True)


if __name__ == "__main__":
    main()
