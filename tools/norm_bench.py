#!/usr/bin/env python
"""Micro-benchmark of the InstanceNorm + ReLU forward / backward streaming kernels at the big 3D layer shapes; prints achieved
GB/s of the algorithmic bytes against the measured HBM peak.  Used under ncu."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icl_b200 import ops  # noqa: E402
from icl_b200.ops import P, c_int, c_ll, call  # noqa: E402


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
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk_path)).get("hbm_gbs", 6546.2) if os.path.exists(pk_path) else 6546.2
    g = torch.Generator(device="cuda").manual_seed(1)
    for B, r, C in ((2, 96, 16), (2, 48, 32), (2, 24, 64)):
        S = r ** 3
        y = torch.randn(B, r, r, r, C, device="cuda", generator=g)
        dA = torch.randn(B, r, r, r, C, device="cuda", generator=g)
        stats = torch.zeros(B, C, 2, dtype=torch.float64, device="cuda")
        call("icl_instnorm_stats", P(y), P(stats), c_int(B), c_int(C), c_ll(S))
        mr = ops.instnorm_finalize(stats, B, C, S)
        n = y.numel()
        pk = ops.empty_pk(B, C, r, r, r, y.device)
        a = torch.empty_like(y)
        # lean forward: fp32 in -> PK (hi + lo) out = 8 B / element; full: + fp32 out = 12 B
        ms = timeit(lambda: call("icl_instnorm_relu_fwd", P(y), P(mr), P(None), P(pk), c_int(1), c_int(B), c_int(C), c_ll(S)))
        print("normact_fwd lean B%d r%d C%d  %7.3f ms  %6.0f GB/s (%4.1f%%)" % (B, r, C, ms, 8e-6 * n / ms, 100 * 8e-6 * n / ms / peak), flush=True)
        ms = timeit(lambda: call("icl_instnorm_relu_fwd", P(y), P(mr), P(a), P(pk), c_int(1), c_int(B), c_int(C), c_ll(S)))
        print("normact_fwd full B%d r%d C%d  %7.3f ms  %6.0f GB/s (%4.1f%%)" % (B, r, C, ms, 12e-6 * n / ms, 100 * 12e-6 * n / ms / peak), flush=True)
        # backward (lean): reduce pass reads dA, y (8 B); apply pass reads dA, y and writes PK (12 B) = 20 B / element
        ms = timeit(lambda: ops.instnorm_relu_bwd(dA, y, mr, True, want_dbias=True, want_f32=False))
        print("normact_bwd lean B%d r%d C%d  %7.3f ms  %6.0f GB/s (%4.1f%%)" % (B, r, C, ms, 20e-6 * n / ms, 100 * 20e-6 * n / ms / peak), flush=True)


if __name__ == "__main__":
    main()
