#!/usr/bin/env python
"""Time one training step (forward + the five ICL losses + backward, no optimizer) of BASELINE config 1 (2D UNet_icl, 12 + 12
slices of 1x256x256) and config 4 (Swin-UNet ICL, 8 + 8 slices of 1x224x224) on cuda:0 with CUDA events."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icl_b200.utils import losses as L  # noqa: E402


def run(name, net, x, y, n_lab, size, iters=5):
    K = 4
    net.cuda().train()
    graph = "--graph" in sys.argv
    from icl_b200.optim import SGD
    opt = SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4, fused_factored=True) if graph else None
    aux, pse = L.AuxLoss(K, resize=[size, size]), L.PseudoSoftLoss(K, resize=[size, size])
    ce_l, dice_l = L.CrossEntropyLoss(), L.DiceLoss(K)

    def step():
        for p in net.parameters():
            p.grad = None
        o = net(x[:n_lab], x[n_lab:])
        loss = ce_l(o[0], y[:n_lab].long()) + dice_l(o[0], y[:n_lab].unsqueeze(1), softmax=True) + aux(o[2], y[:n_lab]) \
            + pse(o[3], o[1]) + 50 * L.softmax_mse_loss(o[3], o[4])
        loss.backward()
        if opt is not None:
            opt.step()
        return loss
    for _ in range(2):
        step()
    if graph:
        from icl_b200.graph import GraphedStep
        eager = step
        try:
            g = GraphedStep(lambda: eager(), (), opt, warmup=1)
            step = lambda: g()   # noqa: E731
            name += " [cuda-graph replay, incl. fused SGD step]"
        except Exception as e:  # report, then time the eager step
            print("%s: graph capture failed: %r" % (name, e), flush=True)
            name += " [eager, incl. fused SGD step]"
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); loss = step(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    if "--kernels" in sys.argv:   # per-kernel device time of the (replayed) step from CUPTI records
        import collections
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(2):
                step()
            torch.cuda.synchronize()
        agg = collections.defaultdict(lambda: [0.0, 0])
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                n = ev.name.replace("(anonymous namespace)::", "").split("(")[0].replace("void ", "")
                agg[n][0] += ev.device_time
                agg[n][1] += 1
        tot = sum(v[0] for v in agg.values())
        print("# %s: %d kernels, %.2f ms kernel time per step" % (name, sum(v[1] for v in agg.values()) // 2, tot / 2e3))
        for n, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
            print("  %-60s %7.1f x %9.1f us/step %5.1f%%  %8.2f us/launch" % (n[:60], c / 2, us / 2, 100 * us / tot, us / c))
    print("%s: %.1f ms/step (fwd + 5 losses + bwd, eager), %.0f slices/s, loss %.4f" % (name, ms, 1e3 * x.shape[0] / ms, loss.item()), flush=True)


def main():
    from icl_b200.networks.unet_icl import UNet_icl
    from icl_b200.networks.vision_transformer import SwinUnet, swin_tiny_lite_config
    g = torch.Generator().manual_seed(1337)
    torch.manual_seed(1337)
    net = UNet_icl(1, 4)  # constructor init (random weights), synthetic inputs
    run("config 1 (UNet_icl, 24 x 1x256x256)", net, torch.randn(24, 1, 256, 256, generator=g).cuda(),
        torch.randint(0, 4, (24, 256, 256), generator=g).cuda(), 12, 256)
    del net
    net = SwinUnet(swin_tiny_lite_config(), img_size=224, num_classes=4)
    run("config 4 (SwinUnet ICL, 16 x 1x224x224)", net, torch.randn(16, 1, 224, 224, generator=g).cuda(),
        torch.randint(0, 4, (16, 224, 224), generator=g).cuda(), 8, 224)


if __name__ == "__main__":
    main()
