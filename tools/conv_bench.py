#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 implicit-GEMM convolution at the unet_3D layer shapes.

    python tools/conv_bench.py [--precision parity|fast] [--only NAME] [--iters N] [--wgrad]

Prints per layer: time per launch (CUDA events on the launching stream, inputs > L2 or L2 flushed between
launches), algorithmic TFLOP/s and the fraction of the measured bf16 peak.  Used under `ncu --set full -k regex:conv3d_umma`."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LAYERS = [  # name, B, cins, cout, r
    ("enc0.conv2", 2, [16], 16, 96), ("enc1.conv1", 2, [16], 32, 48), ("enc1.conv2", 2, [32], 32, 48),
    ("enc2.conv2", 2, [64], 64, 24), ("enc3.conv2", 2, [128], 128, 12), ("center.conv2", 2, [256], 256, 6),
    ("up4.conv1", 2, [128, 256], 128, 12), ("up3.conv1", 2, [64, 128], 64, 24), ("up2.conv1", 2, [32, 64], 32, 48),
    ("up1.conv1", 2, [16, 32], 16, 96), ("up1.conv1.dgrad", 2, [16], 48, 96), ("enc0.conv2.B4", 4, [16], 16, 96),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="parity")
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--wgrad", action="store_true", help="time the weight-gradient kernel (conv3d_wgrad_umma.cu) instead of the forward")
    a = ap.parse_args()
    import icl_b200
    from icl_b200 import ops
    icl_b200.set_precision(a.precision)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = peaks.get("bf16_tflops", 1590.0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, B, cins, cout, r in LAYERS:
        if a.only and a.only != name:
            continue
        g = torch.Generator(device="cuda").manual_seed(1)
        xs = [torch.randn(B, r, r, r, c, device="cuda", generator=g) for c in cins]
        pks = [ops.pack_pk(x) for x in xs]
        w = torch.randn(cout, sum(cins), 3, 3, 3, device="cuda", generator=g) * 0.05
        wp = ops.pack_w_umma(w, False, r)
        bias = torch.zeros(cout, device="cuda")
        stats = torch.zeros(B, cout, 2, dtype=torch.float64, device="cuda")
        out = torch.empty(B, r, r, r, cout, device="cuda")
        if a.wgrad:
            if name.endswith("dgrad") or r % 2:
                continue
            dy_pk = ops.pack_pk(torch.randn(B, r, r, r, cout, device="cuda", generator=g))
            run = lambda: ops.conv3d_wgrad_umma(pks, cins, dy_pk, cout, B, r, r, r)
        else:
            run = lambda: ops.conv3d_umma(pks, cins, wp, bias, cout, B, r, r, r, stats, out=out)
        for _ in range(3):
            run()
        ts = []
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        gf = 2e-9 * 27 * sum(cins) * cout * B * r ** 3
        tf = gf / ms
        print(("wgrad " if a.wgrad else "") + "%-18s B%d r%-3d %-9s->%-3d  %8.3f ms  %7.1f TFLOP/s algorithmic  %5.1f%% of measured bf16 peak (%s)" % (
            name, B, r, "+".join(map(str, cins)), cout, ms, tf, 100 * tf / peak, a.precision), flush=True)


if __name__ == "__main__":
    main()
