"""Batch-sharded data parallelism for the ICL training step — new functionality, the
reference is single-process (every train_*.py calls .cuda() on one device).

One process per GPU (torchrun), each rank runs the reference step on its local (2 labeled + 2 unlabeled)
batch; BatchNorm batch statistics and the proxy mean stay rank-local (DDP-without-SyncBN semantics).
Parity definition: the R-rank gradient equals the mean of the R single-process gradients.

Exchange, per step, all of it on a side stream so that it overlaps the rest of backward:
  * the mlp2 weights (13 824 x 13 824 x 4 tensors = 3.1 GB of gradient) are never all-reduced.  dW = dY^T X has
    rank <= rows (16 per rank at K=2, 128 at K=16), so ranks all-gather the factors (dY, X: rows x N each) the
    moment LinearFn.backward produces them — they are the first thing backward computes — and every rank forms
    the summed, averaged dW locally inside the optimizer's update kernel: >100x less NVLink traffic, same sum
    (different fp32 order).
  * "normal" parameters (8.3 M, 33 MB fp32) travel in buckets (ICL heads / decoder / encoder), each all-reduced
    as soon as the last of its gradients has been accumulated (post-accumulate-grad hooks; how many
    accumulations a parameter receives per step — e.g. two for the backbone, one from the unlabeled and one
    from the labeled pass — is learned on the first step).  The head bucket finishes before the backbone
    backward starts; the backbone is ONE autograd node, so its two buckets complete at the end of backward.
GradAverager.average() joins the side stream and copies the averaged buckets back into .grad.
"""
import torch
import torch.distributed as dist

from . import lanes

FACTOR_MIN_NUMEL = 1 << 24  # weights at least this large are exchanged as factors

_CTX = {"world": 1, "group": None, "factored_ids": set(), "comm": None, "overlap": False}


def factor_context():
    return _CTX


def _comm_stream():
    """Side stream of the exchange (None: exchange on the current stream)."""
    if not _CTX["overlap"] or not torch.cuda.is_available():
        return None
    if _CTX["comm"] is None:
        _CTX["comm"] = torch.cuda.Stream()
    return _CTX["comm"]


def gather_factors(dy2d, x2d, group=None):
    """All-gather the rank-local wgrad factors (blocking form).  Returns (dy_all [R*M, N], x_all [R*M, K])."""
    world = dist.get_world_size(group)
    M = dy2d.shape[0]
    dy_all = torch.empty((world * M, dy2d.shape[1]), dtype=dy2d.dtype, device=dy2d.device)
    x_all = torch.empty((world * M, x2d.shape[1]), dtype=x2d.dtype, device=x2d.device)
    dist.all_gather_into_tensor(dy_all, dy2d.contiguous(), group=group)
    dist.all_gather_into_tensor(x_all, x2d.contiguous(), group=group)
    return dy_all, x_all


def gather_factors_async(dy2d, x2d, group=None):
    """All-gather the factors on the exchange stream.  Returns (dy_all, x_all, event, keepalive): the gathered tensors are
    valid once `event` has been waited on; `keepalive` must stay referenced until then (the inputs are still being read)."""
    comm = _comm_stream()
    if comm is None:
        dy_all, x_all = gather_factors(dy2d, x2d, group)
        return dy_all, x_all, None, None
    world = dist.get_world_size(group)
    M = dy2d.shape[0]
    dy_c, x_c = dy2d.contiguous(), x2d.contiguous()
    dy_all = torch.empty((world * M, dy2d.shape[1]), dtype=dy2d.dtype, device=dy2d.device)
    x_all = torch.empty((world * M, x2d.shape[1]), dtype=x2d.dtype, device=x2d.device)
    comm.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(comm):
        dist.all_gather_into_tensor(dy_all, dy_c, group=group)
        dist.all_gather_into_tensor(x_all, x_c, group=group)
        ev = torch.cuda.Event()
        ev.record(comm)
    return dy_all, x_all, ev, (dy_c, x_c)


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


def default_bucket_of(name):
    """Bucket index by parameter name, in the order backward completes them: ICL heads, decoder, encoder."""
    if name.startswith("sspa.") or name.startswith("uscl."):
        return 0
    if name.startswith("up_concat") or name.startswith("final") or name.startswith("decoder") or name.startswith("out_conv"):
        return 1
    return 2


class _Bucket:
    def __init__(self, idx):
        self.idx = idx
        self.params = []
        self.expected = None   # {param id: accumulations per step}, learned on the first step
        self.count = {}
        self.inflight = None   # (flat, params) once launched
        self.bytes = 0


class GradAverager:
    """Call .begin_step() before the forward pass (optional: it only resets the per-step bookkeeping, .average() does it too),
    and .average() between loss.backward() and optimizer.step()."""

    def __init__(self, model, world_size, group=None, factored=True, overlap=True, bucket_of=default_bucket_of):
        named = list(model.named_parameters())
        self.params = [p for _, p in named]
        self.world, self.group = world_size, group
        _CTX["world"], _CTX["group"], _CTX["overlap"] = world_size, group, bool(overlap)
        _CTX["factored_ids"] = set(id(p) for p in self.params if factored and p.dim() == 2 and p.numel() >= FACTOR_MIN_NUMEL)
        self.overlap = bool(overlap)
        self.buckets = {}
        self._bucket_of = {}
        for name, p in named:
            if id(p) in _CTX["factored_ids"] or getattr(p, "_icl_factors", None) is not None:
                continue
            b = self.buckets.setdefault(bucket_of(name), _Bucket(bucket_of(name)))
            b.params.append(p)
            self._bucket_of[id(p)] = b
        self._hooks = []
        if world_size > 1 and self.overlap:
            for p in self.params:
                if id(p) in self._bucket_of and p.requires_grad:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.bytes_last = 0
        self.launched_in_backward = 0   # buckets whose all-reduce started from a hook during the last backward
        self._launched = 0

    # ------------------------------------------------------------------ per-step bookkeeping
    def begin_step(self):
        for b in self.buckets.values():
            b.count = {}
            b.inflight = None
        self._launched = 0

    def _on_grad(self, p):
        b = self._bucket_of.get(id(p))
        if b is None or b.inflight is not None:
            return
        b.count[id(p)] = b.count.get(id(p), 0) + 1
        if b.expected is not None and b.count == b.expected:
            self._launch(b)
            self._launched += 1

    def _launch(self, b):
        todo = [p for p in b.params if p.grad is not None]
        if not todo:
            b.inflight = (None, [])
            return
        lanes.join_all()   # a bucket's gradients may have been accumulated on different ICL-head lanes
        flat = torch.cat([p.grad.reshape(-1) for p in todo])
        comm = _comm_stream()
        if comm is not None:
            comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(comm):
                dist.all_reduce(flat, group=self.group)
                flat.mul_(1.0 / self.world)
        else:
            dist.all_reduce(flat, group=self.group)
            flat.mul_(1.0 / self.world)
        b.inflight = (flat, todo)
        b.bytes = flat.numel() * 4

    def average(self):
        if self.world <= 1:
            return
        for k in sorted(self.buckets):
            b = self.buckets[k]
            if b.inflight is None:
                self._launch(b)
        comm = _comm_stream()
        if comm is not None:
            torch.cuda.current_stream().wait_stream(comm)   # joins the bucket all-reduces AND the factor all-gathers
        total = 0
        for k in sorted(self.buckets):
            b = self.buckets[k]
            flat, todo = b.inflight
            if flat is None:
                continue
            views, off = [], 0
            for p in todo:
                n = p.numel()
                views.append(flat[off:off + n].view_as(p.grad))
                off += n
            # one multi-tensor copy instead of one small launch per parameter (each is a node of the captured step graph)
            torch._foreach_copy_([p.grad for p in todo], views)
            total += b.bytes
            if self.overlap and b.count:
                b.expected = dict(b.count)   # accumulations per parameter per step (stable from step to step)
        self.bytes_last = total
        self.launched_in_backward = self._launched
        self.begin_step()

    def sync_buffers(self, model, mode="mean"):
        """BatchNorm running statistics are rank-local during training (DDP-without-SyncBN semantics).  Call this before saving a
        checkpoint or switching to eval so that every rank holds the same buffers: floating-point buffers are averaged over ranks
        (mode="mean") or taken from rank 0 (mode="rank0"); integer buffers (num_batches_tracked) always come from rank 0."""
        if self.world <= 1:
            return
        for buf in model.buffers():
            if buf.is_floating_point() and mode == "mean":
                dist.all_reduce(buf, group=self.group)
                buf.mul_(1.0 / self.world)
            else:
                dist.broadcast(buf, src=0, group=self.group)

    def check_consistent(self):
        """Debug aid: every rank must exchange the same bucket sizes (same set of non-None gradients), otherwise the flat
        all-reduce would hang or mix parameters.  One small all-gather; call once after the first backward."""
        sizes = torch.tensor([sum(p.numel() for p in b.params if p.grad is not None) for _, b in sorted(self.buckets.items())],
                             dtype=torch.int64, device=self.params[0].device)
        allsz = [torch.empty_like(sizes) for _ in range(self.world)]
        dist.all_gather(allsz, sizes, group=self.group)
        for r, s in enumerate(allsz):
            if not torch.equal(s, sizes):
                raise RuntimeError("GradAverager: rank %d exchanges bucket sizes %s, this rank %s" % (r, s.tolist(), sizes.tolist()))

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
        _CTX["world"], _CTX["group"], _CTX["factored_ids"], _CTX["overlap"] = 1, None, set(), False
