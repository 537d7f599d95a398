#!/usr/bin/env python
"""Per-kernel device time INSIDE the CUDA-graph replay of the training step (CUPTI activity records via torch.profiler).
The eager CUDA-event breakdown of bench.py times each launch with the host in the loop, which inflates kernels of a few
microseconds; this tool shows where the replayed step really spends its time.

    python tools/graph_kernel_times.py [cfg2|cfg3] [steps]
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
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    a = types.SimpleNamespace(unfused_optimizer=False, no_overlap=False, precision="parity", no_graph=False)
    import icl_b200
    icl_b200.set_precision("parity")
    dev = torch.device("cuda:0")
    tb = bench.TrainBench(a, wl, 0, 1, dev)
    tb.capture()
    for _ in range(3):
        tb.step(tb.x_dev, tb.y_dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tb.step(tb.x_dev, tb.y_dev)
    e1.record()
    torch.cuda.synchronize()
    print("# %s: %.3f ms/step by CUDA events without the profiler" % (wl, e0.elapsed_time(e1) / steps))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            tb.step(tb.x_dev, tb.y_dev)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0.0, 0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            n = ev.name.replace("(anonymous namespace)::", "").split("(")[0].replace("void ", "")
            agg[n][0] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
            agg[n][1] += 1
    tot = sum(v[0] for v in agg.values())
    print("# %d kernel records per step, %.3f ms of kernel time per step" % (sum(v[1] for v in agg.values()) // steps, tot / steps / 1e3))
    print("%-64s %9s %9s %7s %9s" % ("kernel", "launches", "us/step", "share", "us/launch"))
    for n, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("%-64s %9.1f %9.1f %6.1f%% %9.2f" % (n[:64], c / steps, us / steps, 100 * us / tot, us / c))
    tb.close()


if __name__ == "__main__":
    main()
