"""Fused multi-tensor SGD — drop-in for `optim.SGD(model.parameters(), lr, momentum=0.9, weight_decay=1e-4)`
as used by the ICL loops (train_inherent_consistent_unet_3D_BraTS.py:85-86,113-119).

Same update rule as torch.optim.SGD (dampening 0, no nesterov); parameters whose .grad is None are skipped
entirely (no weight decay, no momentum) exactly like torch.  One kernel launch per param group per step
(the 785 M-parameter model is otherwise ~230 launches), lr read from device memory.
"""
import os

import numpy as np
import torch

from . import _lib, lanes, ops
from .ops import P, c_f, c_int, call


class SGD(torch.optim.Optimizer):
    def __init__(self, params, lr=0.01, momentum=0.0, dampening=0, weight_decay=0.0, nesterov=False, fused_factored=False,
                 factored_min_numel=1 << 24):
        """fused_factored=True: 2-D weights with at least `factored_min_numel` elements (the 13 824^2 mlp2 Linears) never get
        a materialised .grad — icl_b200.functional.LinearFn hands their rank-<=rows factors (dY, X) to this optimizer, which
        forms dY^T X inside the update kernel.  Same update rule and numerics as the unfused path up to fp32 summation order."""
        if dampening != 0 or nesterov:
            raise NotImplementedError("icl_b200.optim.SGD implements dampening=0, nesterov=False (the ICL loops' configuration)")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        self._factored = []
        if fused_factored:
            for group in self.param_groups:
                for p in group["params"]:
                    if p.dim() == 2 and p.numel() >= factored_min_numel:
                        p._icl_factors = []
                        self._factored.append(p)
        self._chunk = _lib.lib().icl_sgd_chunk()
        self._cache = {}
        self._lr_dev = {}     # per group: persistent 1-element device tensor the kernels read the learning rate from
        self._tab_cache = {}  # per group: (key, pinned host table, device table) of (param, grad, momentum) pointers

    def _chunks(self, numels, device):
        key = (tuple(numels), str(device))
        hit = self._cache.get(key)
        if hit is None:
            ct, co = [], []
            for i, n in enumerate(numels):
                offs = np.arange(0, n, self._chunk, dtype=np.int64)
                ct.append(np.full(offs.shape, i, dtype=np.int32))
                co.append(offs)
            ct = torch.from_numpy(np.concatenate(ct)).to(device)
            co = torch.from_numpy(np.concatenate(co)).to(device)
            hit = (ct, co)
            self._cache[key] = hit
        return hit

    def zero_grad(self, set_to_none=True):
        for p in self._factored:
            p._icl_factors.clear()
        return super().zero_grad(set_to_none=set_to_none)

    def _step_factored(self, group, lr):
        """The updates run on a lane of their own that waits only for the factors (events recorded where backward produced them):
        inside a captured step the HBM-bound updates then run next to the tensor-core-bound backbone backward instead of after it
        (config 2: 13.4 -> 12.7 ms per step).  While the update is HBM-bound (R <= 256) it gets 64 persistent CTAs, which still pull
        most of the bandwidth and leave the other SMs to the convolutions; tensor-bound updates (large R: data parallel) get 96
        (config 3 with the factor rows of 8 ranks, ICL_SIM_RANKS=8: 19.6 ms with 148 CTAs, 19.0 with 96, 20.0 with 64).  ICL_OPT_LANE=0 disables the lane, ICL_OPT_CTAS overrides the CTA count.  Eager steps are enqueued in program order
        either way."""
        lane = main = None
        ctas_env = -1
        dev = group["params"][0].device
        if os.environ.get("ICL_OPT_LANE", "1") != "0" and lanes.enabled(dev):
            L = lanes.get(dev, ["opt"])
            lane, main = L["opt"], L["main"]
            ctas_env = int(os.environ.get("ICL_OPT_CTAS", "-1"))
            if not torch.cuda.is_current_stream_capturing():
                lane.wait_stream(main)   # eager: the learning-rate refresh and first-step momentum buffers were enqueued on main
        for p in group["params"]:
            fs = getattr(p, "_icl_factors", None)
            if not fs:
                continue
            if p.grad is not None:
                raise RuntimeError("icl_b200.optim.SGD: a fused-factored weight also has a materialised .grad")
            st = self.state[p]
            if "momentum_buffer" not in st:
                st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                if lane is not None:
                    lane.wait_stream(main)
            tgt = lane if lane is not None else torch.cuda.current_stream()
            for f in fs:
                if len(f) > 3 and f[3] is not None:
                    tgt.wait_event(f[3])   # factors produced on an ICL-head lane / all-gathered on the exchange stream (icl_b200.parallel)
                if lane is not None:
                    f[0].record_stream(lane)
                    f[1].record_stream(lane)
            ctas = 0
            if lane is not None:
                ctas = ctas_env if ctas_env >= 0 else (64 if sum(f[0].shape[0] for f in fs) <= 256 else 96)
            with lanes.on(lane):
                ops.sgd_factored(p, st["momentum_buffer"], [(f[0], f[1], f[2] if len(f) > 2 else 1.0) for f in fs], lr, group["momentum"],
                                 group["weight_decay"], max_ctas=ctas)
            fs.clear()
        if lane is not None:
            main.wait_stream(lane)

    def _lr_tensor(self, gi, group, dev):
        """Learning rate in device memory.  Refreshed from param_groups on every eager step; inside CUDA-graph capture the
        refresh is skipped (it would bake the value in) — call sync_lr() before each replay instead."""
        t = self._lr_dev.get(gi)
        if t is None or t.device != dev:
            t = torch.full((1,), float(group["lr"]), dtype=torch.float32, device=dev)
            self._lr_dev[gi] = t
        elif not torch.cuda.is_current_stream_capturing():
            t.fill_(float(group["lr"]))
        return t

    def sync_lr(self):
        for gi, group in enumerate(self.param_groups):
            t = self._lr_dev.get(gi)
            if t is not None:
                t.fill_(float(group["lr"]))

    def _table(self, gi, rows, dev):
        """Device table of (p, g, m, n) rows.  Gradient tensors are re-allocated every step, but the caching allocator (and a
        CUDA graph's private pool) returns the same addresses, so a table is uploaded only when a pointer set is new — from
        pinned memory, asynchronously.  Tables are never overwritten (a captured graph keeps replaying its upload), and one
        spare pinned/device pair is kept ready because pinned memory cannot be allocated during stream capture."""
        key = tuple(rows)
        cache = self._tab_cache.setdefault(gi, {"tabs": {}, "spare": None})
        captured = cache.setdefault("captured", {})
        hit = captured.get(key) or cache["tabs"].get(key)
        capturing = torch.cuda.is_current_stream_capturing()
        if hit is not None and capturing:
            captured[key] = hit
        if hit is None:
            shape = (len(rows), 4)
            sp = cache["spare"]
            if sp is not None and tuple(sp[0].shape) == shape and sp[1].device == dev:
                host, tab = sp
                cache["spare"] = None
            elif capturing:
                raise RuntimeError("icl_b200.optim.SGD: run at least one eager step before capturing a CUDA graph")
            else:
                host = torch.empty(shape, dtype=torch.int64).pin_memory()
                tab = torch.empty(shape, dtype=torch.int64, device=dev)
            host.numpy()[...] = np.asarray(rows, dtype=np.int64)
            tab.copy_(host, non_blocking=True)
            if capturing:
                # a captured graph replays this upload for as long as it lives: its table is never evicted
                hit = cache["captured"][key] = (host, tab)
            else:
                if len(cache["tabs"]) >= 8:
                    cache["tabs"].pop(next(iter(cache["tabs"])))
                hit = cache["tabs"][key] = (host, tab)
        if cache["spare"] is None and not capturing:
            shape = (len(rows), 4)
            cache["spare"] = (torch.empty(shape, dtype=torch.int64).pin_memory(), torch.empty(shape, dtype=torch.int64, device=dev))
        return hit[1]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lanes.join_all()   # gradients / factor lists produced on the ICL-head lanes (no-op when no lane exists)
        for gi, group in enumerate(self.param_groups):
            rows, keep = [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("icl_b200.optim.SGD needs contiguous float32 CUDA parameters (no CPU fallback)")
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                st = self.state[p]
                if "momentum_buffer" not in st:
                    st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                m = st["momentum_buffer"]
                rows.append((p.data_ptr(), g.data_ptr(), m.data_ptr(), p.numel()))
                keep.append(g)
            dev = group["params"][0].device
            lr = self._lr_tensor(gi, group, dev)
            if self._factored:
                with torch.cuda.device(dev):
                    self._step_factored(group, lr)
            if not rows:
                continue
            tab = self._table(gi, rows, dev)
            ct, co = self._chunks([r[3] for r in rows], dev)
            with torch.cuda.device(dev):
                call("icl_sgd_multi", P(tab), P(ct), P(co), c_int(ct.numel()), P(lr), c_f(group["momentum"]), c_f(group["weight_decay"]),
                     c_int(0), mbytes=20e-6 * sum(r[3] for r in rows))
        ops.weights_changed()   # packed tensor-core weight operands are stale now (the kernels wrote through raw pointers)
        return loss
