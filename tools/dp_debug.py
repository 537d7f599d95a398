#!/usr/bin/env python
"""Diagnostic for the 2-rank data-parallel step (run under torchrun --nproc-per-node 2): compares DP gradients against the mean of the
rank-local single-process gradients, per parameter, for overlap on / off, and two single-process runs against each other (the
run-to-run noise floor)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_dp import _build, _loss  # noqa: E402
from icl_b200 import parallel  # noqa: E402
from icl_b200.utils import synth  # noqa: E402


def grads_of(net):
    return {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in net.named_parameters()}


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    K = 2
    weights = (1.0, 1.0, 1.0, 1.0, 10.0)
    xs = [synth.synth_volume((4, 1, 96, 96, 96), 1338 + r).to(dev) for r in range(world)]
    ys = [synth.synth_labels((4, 96, 96, 96), K, 1339 + r).to(dev) for r in range(world)]
    net = _build(K, dev)
    # single-process gradients of every rank's batch, twice (noise floor)
    local = []
    for rep in range(2):
        per = []
        for r in range(world):
            net.zero_grad(set_to_none=True)
            _loss(net, xs[r], ys[r], K, weights).backward()
            per.append(grads_of(net))
        local.append(per)
    mean = {k: (None if local[0][0][k] is None else sum(local[0][r][k] for r in range(world)) / world) for k in local[0][0]}

    def report(tag, got, ref, top=6):
        rows = []
        for k in ref:
            if ref[k] is None:
                assert got[k] is None, k
                continue
            n = ref[k].double().norm().item()
            if n < 1e-5:   # mathematically-zero gradients (biases in front of a normalisation): rounding noise only
                continue
            rows.append(((got[k] - ref[k]).double().norm().item() / max(n, 1e-30), k, n))
        rows.sort(reverse=True)
        if rank == 0:
            print(tag, " | ".join("%s %.2e (|g| %.1e)" % (k, e, n) for e, k, n in rows[:top]), flush=True)

    report("noise floor (same batch, two runs):", local[1][0], local[0][0])
    for overlap in (False, True):
        dp = parallel.GradAverager(net, world, factored=True, overlap=overlap)
        for step in range(3):
            dp.begin_step()
            net.zero_grad(set_to_none=True)
            _loss(net, xs[rank], ys[rank], K, weights).backward()
            dp.average()
            torch.cuda.synchronize()
            report("overlap=%s step %d launched_in_backward=%d:" % (overlap, step, dp.launched_in_backward), grads_of(net), mean)
        dp.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
