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
    rows_only = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # restrict to one row count (ncu captures)
    r_only = int(sys.argv[3]) if len(sys.argv) > 3 else 0         # restrict the sgd sweep to one R
    g = torch.Generator(device="cuda").manual_seed(1)
    for rows, n in ((16, 13824), (128, 13824), (32, 1728), (256, 1728)):
        if rows_only and rows != rows_only:
            continue
        x = torch.randn(rows, n, device="cuda", generator=g)
        w = torch.randn(n, n, device="cuda", generator=g) * 0.01
        b = torch.zeros(n, device="cuda")
        mb = n * n * 4 / 1e6
        for mode in ("umma", "legacy"):
            os.environ["ICL_DISABLE_BIGW"] = "0" if mode == "umma" else "1"
            if only in ("", "fwd"):
                ms = timeit(lambda: ops.linear_fwd(x, w, b, 1, want_pre=True))
                print("%-6s linear_fwd   %4dx%dx%d  %8.3f ms  %7.0f GB/s (%4.1f%% of %.0f)  %6.1f TFLOP/s" % (
                    mode, rows, n, n, ms, mb / ms, 100 * mb / ms / peak, peak, 2e-9 * rows * n * n / ms), flush=True)
            if only in ("", "dgrad"):
                ms = timeit(lambda: ops.linear_dgrad(x, w))
                print("%-6s linear_dgrad %4dx%dx%d  %8.3f ms  %7.0f GB/s (%4.1f%%)" % (mode, rows, n, n, ms, mb / ms, 100 * mb / ms / peak), flush=True)
        os.environ["ICL_DISABLE_BIGW"] = "0"
        if only in ("", "sgd") and n == 13824 and rows == 16:
            m = torch.zeros_like(w)
            lr = torch.full((1,), 0.01, device="cuda")
            for R in (32, 64, 256, 512, 1024, 2048):
                if r_only and R != r_only:
                    continue
                dy = torch.randn(R, n, device="cuda", generator=g) * 0.01
                xx = torch.randn(R, n, device="cuda", generator=g)
                for mode in ("umma", "legacy"):
                    if mode == "legacy" and R > 256:
                        continue
                    os.environ["ICL_DISABLE_BIGW"] = "0" if mode == "umma" else "1"
                    ms = timeit(lambda: ops.sgd_factored(w, m, [(dy, xx, 1.0)], lr, 0.9, 1e-4), iters=4)
                    print("%-6s sgd_factored R%-4d %dx%d  %8.3f ms  %7.0f GB/s of 16 B/param (%4.1f%%)  %6.1f TFLOP/s" % (
                        mode, R, n, n, ms, 4 * mb / ms, 100 * 4 * mb / ms / peak, 2e-9 * R * n * n / ms), flush=True)
            os.environ["ICL_DISABLE_BIGW"] = "0"


if __name__ == "__main__":
    main()
