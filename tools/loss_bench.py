#!/usr/bin/env python
"""Micro-benchmark of the loss kernels (class statistics family, softmax-MSE) at the BASELINE label grid 2 x 96^3, K classes:
forward and backward of each of the five ICL loss terms, CUDA-event time and achieved GB/s of the algorithmic bytes (SURVEY.md
section 8d).  Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel launch list."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icl_b200.utils import losses as L  # noqa: E402
from icl_b200.utils import synth  # noqa: E402


def timeit(fn, iters=5):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
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
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk)).get("hbm_gbs", 6546.2) if os.path.exists(pk) else 6546.2
    B, V = 2, 96 ** 3
    labels = synth.synth_labels((B, 96, 96, 96), K, 31).cuda()
    cl = lambda t: t.cuda().contiguous(memory_format=torch.channels_last_3d)
    final_lab = cl(synth.synth_volume((B, K, 96, 96, 96), 32)).requires_grad_(True)
    final_unlab = cl(synth.synth_volume((B, K, 96, 96, 96), 33))
    mk = lambda s: [synth.synth_volume((B, K, r, r, r), s + i).cuda().requires_grad_(True) for i, r in enumerate((6, 12, 24))]
    fms, fms2, fms3 = mk(40), mk(50), [t.detach() for t in mk(60)]
    aux, pse = L.AuxLoss3D(K), L.PseudoSoftLoss3D(K)
    cases = {
        "seg ce+dice (main)": (lambda: sum(L.seg_ce_dice(final_lab, labels)), [final_lab], B * V * (4 * K + 8), B * V * (8 * K + 8)),
        "aux (3 scales)": (lambda: aux(fms, labels), fms, 3 * B * V * 8, 3 * B * V * 8),
        "pseudo-soft (3 scales)": (lambda: pse(fms2, final_unlab), fms2, 3 * B * V * 4 * K, 3 * B * V * 4 * K),
        "softmax-mse (3 scales)": (lambda: L.softmax_mse_loss(fms2, fms3), fms2, 0, 0),
    }
    for name, (fn, leaves, fb, bb) in cases.items():
        t_f = timeit(lambda: fn())

        def fb_():
            for t in leaves:
                t.grad = None
            fn().backward()
        t_fb = timeit(fb_)
        t_b = t_fb - t_f
        print("K=%d %-24s fwd %7.3f ms (%6.0f GB/s, %4.1f%% of %.0f)   bwd %7.3f ms (%6.0f GB/s, %4.1f%%)" % (
            K, name, t_f, fb / 1e6 / t_f, 100 * fb / 1e6 / t_f / peak, peak, t_b, bb / 1e6 / max(t_b, 1e-6), 100 * bb / 1e6 / max(t_b, 1e-6) / peak),
            flush=True)


if __name__ == "__main__":
    main()
