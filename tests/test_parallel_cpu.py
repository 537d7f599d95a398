"""CPU, world_size-2 gloo: the data-parallel exchange of icl_b200/parallel.py — bucketed all-reduce of the normal
gradients and the factor all-gather that replaces the 3 GB mlp2 gradient all-reduce.  Parity
definition: the R-rank result equals the mean of the R single-process gradients."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from icl_b200 import parallel
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.Linear(8, 4))
    unused = torch.nn.Parameter(torch.zeros(3))
    model.register_parameter("unused", unused)
    g = torch.Generator().manual_seed(100 + rank)
    for p in model.parameters():
        if p is not unused:
            p.grad = torch.randn(p.shape, generator=g)
    avg = parallel.GradAverager(model, world)
    avg.average()
    grads = {k: (None if p.grad is None else p.grad.clone()) for k, p in model.named_parameters()}
    # factor exchange: dW_mean == mean_r dy_r^T x_r
    dy = torch.randn(5, 6, generator=g)
    x = torch.randn(5, 7, generator=g)

    def wgrad(dy_, x_, acc):
        d = dy_.t() @ x_
        return d if acc is None else acc + d
    dW = parallel.averaged_factored_wgrad(dy, x, wgrad)
    # numpy payloads travel by value: a tensor would be shared through a file descriptor served by this process, which may have
    # exited before the parent receives it
    q.put((rank, {k: (None if v is None else v.numpy()) for k, v in grads.items()}, dy.numpy(), x.numpy(), dW.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_averager_and_factor_exchange_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    tt = lambda a: None if a is None else torch.from_numpy(a)
    res = [(r[0], {k: tt(v) for k, v in r[1].items()}, tt(r[2]), tt(r[3]), tt(r[4])) for r in res]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # expected: mean of the two ranks' local gradients
    exp = []
    for rank in range(world):
        g = torch.Generator().manual_seed(100 + rank)
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.Linear(8, 4))
        model.register_parameter("unused", torch.nn.Parameter(torch.zeros(3)))
        exp.append({k: torch.randn(p.shape, generator=g) for k, p in model.named_parameters() if k != "unused"})
    for rank in range(world):
        grads = res[rank][1]
        assert grads["unused"] is None  # a parameter without gradient stays None on every rank
        for k in exp[0]:
            assert torch.allclose(grads[k], (exp[0][k] + exp[1][k]) / 2, atol=1e-6), k
    want_dW = sum(r[2].t() @ r[3] for r in res) / world
    for r in res:
        assert torch.allclose(r[4], want_dW, atol=1e-5)


# ------------------------------------------------------------------------------------------ window-sharded inference
def _tiny_net():
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Conv3d(1, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv3d(4, 3, 1))
    return net.eval()


def _install_inference_standins(inference, ops):
    """torch stand-ins for the two sliding-window kernels and the CUDA-only guards (host logic only; the kernels are GPU-tested)."""
    def accumulate(logits_ndhwc, score, cnt, xs, ys, zs):
        pw, ph, pd, K = logits_ndhwc.shape
        score[:, xs:xs + pw, ys:ys + ph, zs:zs + pd] += logits_ndhwc.permute(3, 0, 1, 2)
        cnt[xs:xs + pw, ys:ys + ph, zs:zs + pd] += 1

    def finalize(score, cnt):
        return torch.argmax(score / cnt.unsqueeze(0), dim=0)
    inference._accumulate, inference._finalize = accumulate, finalize
    inference._check_device = lambda dev: None
    ops._require_cuda = lambda t: None


def _infer_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from icl_b200 import inference, ops
    _install_inference_standins(inference, ops)
    g = torch.Generator().manual_seed(5)
    image = torch.randn(40, 28, 36, generator=g).numpy()
    label = inference.test_single_case_sharded(_tiny_net(), image, 8, 8, (16, 16, 16), num_classes=3)
    score, cnt, _ = inference.sliding_window_scores(_tiny_net(), image, 8, 8, (16, 16, 16), 3, rank=rank, world_size=world)
    q.put((rank, label, int((cnt > 0).sum()), float(cnt.sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_window_sharded_inference_gloo():
    """Two ranks each evaluate half of the windows; after the all-reduce both hold the label map of the single-process driver."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_infer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    from icl_b200 import inference, ops
    saved = (inference._accumulate, inference._finalize, inference._check_device, ops._require_cuda)
    try:
        _install_inference_standins(inference, ops)
        g = torch.Generator().manual_seed(5)
        image = torch.randn(40, 28, 36, generator=g).numpy()
        want = inference.test_single_case(_tiny_net(), image, 8, 8, (16, 16, 16), num_classes=3)
        _, cnt_all, _ = inference.sliding_window_scores(_tiny_net(), image, 8, 8, (16, 16, 16), 3)
    finally:
        inference._accumulate, inference._finalize, inference._check_device, ops._require_cuda = saved
    assert want.shape == (40, 28, 36)
    for rank, label, covered, visits in res:
        agree = (label == want).mean()
        assert agree >= 0.9999, agree           # identical up to fp32 re-association of the window sums at exact near-ties
    # the two shares partition the windows: visit counts add up to the single-process counts
    assert abs(res[0][3] + res[1][3] - float(cnt_all.sum())) < 1e-3
    assert res[0][3] > 0 and res[1][3] > 0
