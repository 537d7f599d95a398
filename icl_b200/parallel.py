"""Batch-sharded data parallelism for the ICL training step — new functionality, the
reference is single-process (every train_*.py calls .cuda() on one device).

One process per GPU (torchrun), each rank runs the reference step on its local (2 labeled + 2 unlabeled)
batch; BatchNorm batch statistics and the proxy mean stay rank-local (DDP-without-SyncBN semantics).
Parity definition: the R-rank gradient equals the mean of the R single-process gradients.

Exchange, per step:
  * "normal" parameters (8.3 M, 33 MB fp32): one flat bucket, one NCCL all-reduce (NVLS/ring over NVSwitch).
  * the mlp2 weights (13 824 x 13 824 x 4 tensors = 3.1 GB of gradient): never all-reduced.  dW = dY^T X has
    rank <= rows (16 per rank at K=2), so ranks all-gather the factors (dY, X: rows x N each, < 1 MB) and each
    forms the summed, averaged dW locally — >1000x less NVLink traffic, same sum (different fp32 order).
"""
import torch
import torch.distributed as dist

FACTOR_MIN_NUMEL = 1 << 24  # weights at least this large are exchanged as factors

_CTX = {"world": 1, "group": None, "factored_ids": set()}


def factor_context():
    return _CTX


def gather_factors(dy2d, x2d, group=None):
    """All-gather the rank-local wgrad factors.  Returns (dy_all [R*M, N], x_all [R*M, K])."""
    world = dist.get_world_size(group)
    M = dy2d.shape[0]
    packed = torch.cat([dy2d.reshape(-1), x2d.reshape(-1)])
    out = torch.empty((world * packed.numel(),), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    out = out.view(world, packed.numel())
    n_dy = dy2d.numel()
    dy_all = out[:, :n_dy].reshape(world * M, dy2d.shape[1])
    x_all = out[:, n_dy:].reshape(world * M, x2d.shape[1])
    return dy_all.contiguous(), x_all.contiguous()


def averaged_factored_wgrad(dy2d, x2d, wgrad_fn, group=None):
    """mean over ranks of (dy_r^T x_r), computed from gathered factors.  wgrad_fn(dy, x, accumulate) -> dW
    must compute dy^T x for up to 64 rows at a time (icl_outer_wgrad on the GPU)."""
    world = dist.get_world_size(group)
    dy_all, x_all = gather_factors(dy2d, x2d, group)
    dy_all = dy_all * (1.0 / world)
    dW = None
    for r0 in range(0, dy_all.shape[0], 64):
        dW = wgrad_fn(dy_all[r0:r0 + 64].contiguous(), x_all[r0:r0 + 64].contiguous(), dW)
    return dW


class GradAverager:
    """Call .average() between loss.backward() and optimizer.step()."""

    def __init__(self, model, world_size, group=None, factored=True):
        self.params = [p for p in model.parameters()]
        self.world, self.group = world_size, group
        _CTX["world"], _CTX["group"] = world_size, group
        _CTX["factored_ids"] = set(id(p) for p in self.params if factored and p.dim() == 2 and p.numel() >= FACTOR_MIN_NUMEL)
        self.bytes_last = 0

    def average(self):
        if self.world <= 1:
            return
        todo = [p for p in self.params if p.grad is not None and id(p) not in _CTX["factored_ids"]]
        if not todo:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in todo])
        dist.all_reduce(flat, group=self.group)
        flat.mul_(1.0 / self.world)
        views, off = [], 0
        for p in todo:
            n = p.numel()
            views.append(flat[off:off + n].view_as(p.grad))
            off += n
        # one multi-tensor copy instead of ~330 small launches (each is a node of the captured step graph)
        torch._foreach_copy_([p.grad for p in todo], views)
        self.bytes_last = flat.numel() * 4

    def close(self):
        _CTX["world"], _CTX["group"], _CTX["factored_ids"] = 1, None, set()
