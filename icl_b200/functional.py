"""torch.autograd.Functions over the ICL-head kernels.

Each Function is the CUDA replacement of one torch op class the reference dispatches inside
InherentConsistent / Class_Decoder / Query_Attention / SeparableConv3d (networks/unet_3D_icl.py:155-345).
Autograd composes them, so the reference's gradient pruning (dead uscl branches, detached targets) and its
`.grad is None` set fall out of the graph structure exactly as they do for the reference.
"""
import os

import torch

from . import ops, parallel
from .ops import P, c_f, c_int, c_ll, call


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


class LinearFn(torch.autograd.Function):
    """y = act(x @ w.T + b) over the last axis (nn.Linear; also 1x1x1 Conv3d on NDHWC rows and Conv1d k=1)."""

    @staticmethod
    def forward(ctx, x, w, b, act):
        ops._require_cuda(x)
        x2 = _c(x.detach()).reshape(-1, x.shape[-1])
        w_ = _c(w.detach())
        y, pre = ops.linear_fwd(x2, w_, None if b is None else _c(b.detach()), act, want_pre=bool(act))
        ctx.save_for_backward(x2, w_, pre if act else None)
        ctx.has_bias, ctx.act, ctx.xshape, ctx.w_id = b is not None, act, x.shape, id(w)
        ctx.w_ref = w  # the Parameter itself: icl_b200.optim.SGD(fused_factored=True) hangs a factor sink on it
        return y.reshape(x.shape[:-1] + (w.shape[0],))

    @staticmethod
    def backward(ctx, dy):
        x2, w_, pre = ctx.saved_tensors
        g = _c(dy).reshape(-1, w_.shape[0])
        if ctx.act:
            g = ops.gelu_bwd(g, pre)
        dx = ops.linear_dgrad(g, w_).reshape(ctx.xshape) if ctx.needs_input_grad[0] else None
        dw = db = None
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dp = parallel.factor_context()
            sink = getattr(ctx.w_ref, "_icl_factors", None)
            if sink is not None:
                # fused optimizer path: hand the rank-<=rows factors to the optimizer, which forms
                # dy^T x inside the momentum-SGD update; the 764 MB gradient tensor is never materialised (.grad stays None)
                if dp["world"] > 1:
                    # the factors are ready at the very start of backward: gather them on the exchange stream while the rest of
                    # backward runs; the optimizer waits on the event before it packs them
                    gd, xd, ev, keep = parallel.gather_factors_async(g, x2, dp["group"])
                    sink.append((gd, xd, 1.0 / dp["world"], ev, keep))
                else:
                    # the event lets the optimizer start this weight's update as soon as its last factor pair exists (optim.SGD
                    # may run the update on a lane of its own, next to the rest of backward)
                    ev = None
                    sim = int(os.environ.get("ICL_SIM_RANKS", "1"))
                    if sim > 1:
                        # measurement knob: the compute load of `sim` data-parallel ranks on one GPU — the factor rows repeated `sim`
                        # times with weight 1/sim (the mean over `sim` identical ranks: the same gradient, sim-fold rank)
                        g, x2 = g.repeat(sim, 1), x2.repeat(sim, 1)
                    if g.is_cuda:
                        ev = torch.cuda.Event()
                        ev.record()
                    sink.append((g, x2, 1.0 / sim, ev, None))
                if ctx.has_bias:
                    db = torch.empty((w_.shape[0],), dtype=torch.float32, device=g.device)
                    call("icl_colsum", P(g), P(db), c_ll(g.shape[0]), c_int(w_.shape[0]), c_int(0))
            elif dp["world"] > 1 and ctx.w_id in dp["factored_ids"]:
                # data parallel: exchange the rank-<=rows factors instead of all-reducing the 764 MB dW
                dw = parallel.averaged_factored_wgrad(g, x2, ops.outer_wgrad_acc, dp["group"])
                if ctx.has_bias:
                    db = torch.empty((w_.shape[0],), dtype=torch.float32, device=g.device)
                    call("icl_colsum", P(g), P(db), c_ll(g.shape[0]), c_int(w_.shape[0]), c_int(0), tag="%dx%d" % (g.shape[0], w_.shape[0]))
            else:
                dw, db = ops.linear_wgrad(g, x2, ctx.has_bias)
        return dx, dw, db, None


def linear(x, w, b=None, act=0):
    return LinearFn.apply(x, w, b, act)


class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm over the last axis (C, or the spatial axis N for norm3; unet_3D_icl.py:187,248-258)."""

    @staticmethod
    def forward(ctx, x, w, b, eps):
        ops._require_cuda(x)
        C = x.shape[-1]
        x2 = _c(x.detach()).reshape(-1, C)
        rows = x2.shape[0]
        y = torch.empty_like(x2)
        mr = torch.empty((rows, 2), dtype=torch.float32, device=x.device)
        call("icl_layernorm_fwd", P(x2), P(w.detach()), P(b.detach()), P(y), P(mr), c_ll(rows), c_int(C), c_f(eps), mbytes=8e-6 * rows * C,
             tag="%dx%d" % (rows, C))
        ctx.save_for_backward(x2, w.detach(), mr)
        ctx.xshape = x.shape
        return y.reshape(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, w, mr = ctx.saved_tensors
        C = x2.shape[1]
        g = _c(dy).reshape(-1, C)
        dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        want_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dw = ops.zeros((C,), torch.float32, g.device) if want_w else None
        db = ops.zeros((C,), torch.float32, g.device) if want_w else None
        call("icl_layernorm_bwd", P(g), P(x2), P(w), P(mr), P(dx), P(dw), P(db), c_ll(x2.shape[0]), c_int(C), mbytes=12e-6 * x2.numel(),
             tag="%dx%d" % (x2.shape[0], C))
        return (dx.reshape(ctx.xshape) if dx is not None else None), dw, db, None


def layer_norm(x, w, b, eps=1e-5):
    return LayerNormFn.apply(x, w, b, eps)


class ProxyAttnFn(torch.autograd.Function):
    """Query_Attention core (unet_3D_icl.py:287-296).  ql [B,K,C] is read flat as [B,H,K,hd]; kv [B,N,2C].
    Returns (xv [B,K,C] flat view of softmax(map) @ v, map [B,K,H,N] = scaled logits BEFORE softmax)."""

    @staticmethod
    def forward(ctx, ql, kv, H, want_xv):
        ops._require_cuda(ql)
        ql_, kv_ = _c(ql.detach()), _c(kv.detach())
        B, K, C = ql_.shape
        N = kv_.shape[1]
        scale = float(C // H) ** -0.5
        amap = torch.empty((B, K, H, N), dtype=torch.float32, device=ql.device)
        xv = torch.empty((B, K, C), dtype=torch.float32, device=ql.device) if want_xv else None
        mstat = torch.empty((B * H * K, 2), dtype=torch.float32, device=ql.device) if want_xv else None
        call("icl_proxy_attn_fwd", P(ql_), P(kv_), P(amap), P(xv), P(mstat), c_int(B), c_int(N), c_int(C), c_int(H), c_int(K),
             c_f(scale), c_int(1 if want_xv else 0), P(ops.reduce_ws(ql.device)), mbytes=4e-6 * (B * N * C * (2 if want_xv else 1) + B * K * H * N),
             tag="B%d N%d C%d H%d K%d" % (B, N, C, H, K))
        ctx.save_for_backward(ql_, kv_, amap, mstat, xv)
        ctx.dims = (B, N, C, H, K, scale)
        ctx.set_materialize_grads(False)
        return xv, amap

    @staticmethod
    def backward(ctx, dxv, dmap):
        ql_, kv_, amap, mstat, xv = ctx.saved_tensors
        B, N, C, H, K, scale = ctx.dims
        if dxv is None and dmap is None:
            return None, None, None, None
        dl = torch.empty_like(amap)
        dql = torch.empty_like(ql_)
        dkv = torch.empty_like(kv_)
        dmap_c = None if dmap is None else _c(dmap)   # keep the (possibly fresh) contiguous copies alive across the launch
        dxv_c = None if dxv is None else _c(dxv)
        call("icl_proxy_attn_bwd", P(dmap_c), P(dxv_c), P(xv), P(amap), P(ql_), P(kv_),
             P(mstat), P(dl), P(dql), P(dkv), c_int(B), c_int(N), c_int(C), c_int(H), c_int(K), c_f(scale), P(ops.reduce_ws(ql_.device)),
             mbytes=4e-6 * (2 * B * N * 2 * C + 3 * B * K * H * N), tag="B%d N%d C%d H%d K%d" % (B, N, C, H, K))
        return dql, dkv, None, None


def proxy_attention(ql, kv, num_heads, want_xv=True):
    return ProxyAttnFn.apply(ql, kv, num_heads, want_xv)


class WindowAttnFn(torch.autograd.Function):
    """(Shifted-)window multi-head self-attention on token-major tensors (swinunet_icl.py:120-155 + the roll / partition /
    reverse / mask of :249-293): qkv [B, H*W, 3C], relative-position table [(2ws-1)^2, nH] -> [B, H*W, C].
    Probabilities are recomputed in backward; nothing but qkv is saved."""

    @staticmethod
    def forward(ctx, qkv, table, H, W, nH, ws, shift):
        ops._require_cuda(qkv)
        q_, t_ = _c(qkv.detach()), _c(table.detach())
        B, L, C3 = q_.shape
        if L != H * W or C3 % 3:
            raise RuntimeError("window_attention: qkv %s does not match %dx%d tokens" % (tuple(q_.shape), H, W))
        C = C3 // 3
        out = torch.empty((B, L, C), dtype=torch.float32, device=q_.device)
        call("icl_window_attn_fwd", P(q_), P(t_), P(out), c_int(B), c_int(H), c_int(W), c_int(C), c_int(nH), c_int(ws), c_int(shift))
        ctx.save_for_backward(q_, t_)
        ctx.dims = (B, H, W, C, nH, ws, shift)
        return out

    @staticmethod
    def backward(ctx, dout):
        q_, t_ = ctx.saved_tensors
        B, H, W, C, nH, ws, shift = ctx.dims
        dqkv = torch.empty_like(q_)
        dtable = torch.zeros_like(t_)
        dout_c = _c(dout)   # keep the contiguous copy alive across the launch
        call("icl_window_attn_bwd", P(q_), P(t_), P(dout_c), P(dqkv), P(dtable), c_int(B), c_int(H), c_int(W), c_int(C), c_int(nH),
             c_int(ws), c_int(shift))
        return dqkv, dtable, None, None, None, None, None


def window_attention(qkv, table, H, W, num_heads, window_size, shift):
    return WindowAttnFn.apply(qkv, table, H, W, num_heads, window_size, shift)


class AddScaledFn(torch.autograd.Function):
    """out = a + b * r  with r broadcast per leading-axis sample (DropPath residuals, unet_3D_icl.py:264-267);
    r is None in eval mode."""

    @staticmethod
    def forward(ctx, a, b, r):
        a_, b_ = _c(a.detach()), _c(b.detach())
        rows = a_.shape[0]
        out = ops.row_combine(a_, None, b_, r, rows)
        ctx.r, ctx.rows = r, rows
        return out

    @staticmethod
    def backward(ctx, g):
        g = _c(g)
        db = g if ctx.r is None else ops.row_combine(g, ctx.r, None, None, ctx.rows)
        return g, db, None


def add_scaled(a, b, r=None):
    return AddScaledFn.apply(a, b, r)


class BatchMeanFn(torch.autograd.Function):
    """x.mean(dim=0, keepdim=True) (updated_guided_Q.mean, unet_3D_icl.py:224)."""

    @staticmethod
    def forward(ctx, x):
        x_ = _c(x.detach())
        B = x_.shape[0]
        n = x_.numel() // B
        out = torch.empty((1,) + tuple(x_.shape[1:]), dtype=torch.float32, device=x.device)
        ones = torch.full((B,), 1.0 / B, dtype=torch.float32, device=x.device)
        ops.sgemm(1, n, B, ones, B, 1, x_, n, 1, out, n, 1)
        ctx.B = B
        return out

    @staticmethod
    def backward(ctx, g):
        return (g / ctx.B).expand((ctx.B,) + tuple(g.shape[1:]))


def batch_mean(x):
    return BatchMeanFn.apply(x)


class DwConv3dFn(torch.autograd.Function):
    """Depthwise 3x3x3 conv, groups = channels, no bias, on planar [NB, CH, d, h, w] (unet_3D_icl.py:321-323)."""

    @staticmethod
    def forward(ctx, x, w):
        x_, w_ = _c(x.detach()), _c(w.detach())
        NB, CH, d, h, wd = x_.shape
        y = torch.empty_like(x_)
        call("icl_dwconv3d", P(x_), P(w_), P(y), c_int(NB), c_int(CH), c_int(d), c_int(h), c_int(wd), c_int(0))
        ctx.save_for_backward(x_, w_)
        return y

    @staticmethod
    def backward(ctx, dy):
        x_, w_ = ctx.saved_tensors
        NB, CH, d, h, wd = x_.shape
        dy = _c(dy)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x_)
            call("icl_dwconv3d", P(dy), P(w_), P(dx), c_int(NB), c_int(CH), c_int(d), c_int(h), c_int(wd), c_int(1))
        dw = torch.empty_like(w_)
        call("icl_dwconv3d_wgrad", P(x_), P(dy), P(dw), c_int(NB), c_int(CH), c_int(d), c_int(h), c_int(wd), P(ops.reduce_ws(x_.device)),
             mbytes=8e-6 * x_.numel(), tag="NB%d CH%d r%d" % (NB, CH, d))
        return dx, dw


def dwconv3d(x, w):
    return DwConv3dFn.apply(x, w)


class BnReluFn(torch.autograd.Function):
    """BatchNorm3d (training: batch statistics, running-stat update) + ReLU on planar [NB, CH, ...]
    (bn_depth/relu1 and bn_point/relu2 of SeparableConv3d, unet_3D_icl.py:337-342)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, run_mean, run_var, training, momentum, eps):
        x_ = _c(x.detach())
        NB, CH = x_.shape[0], x_.shape[1]
        S = x_.numel() // (NB * CH)
        y = torch.empty_like(x_)
        mr = torch.empty((CH, 2), dtype=torch.float32, device=x.device)
        if training:
            call("icl_bn_relu_fwd", P(x_), P(gamma.detach()), P(beta.detach()), P(y), P(mr), P(run_mean), P(run_var), c_int(NB), c_int(CH),
                 c_ll(S), c_f(eps), c_f(momentum), P(ops.reduce_ws(x_.device)), mbytes=12e-6 * x_.numel(), tag="NB%d CH%d S%d" % (NB, CH, S))
        else:
            raise RuntimeError("icl_b200 BatchNorm: eval-mode ICL heads are not on the reference's path "
                               "(inference returns before the heads, unet_3D_icl.py:119-120)")
        ctx.save_for_backward(x_, y, mr, gamma.detach())
        ctx.dims = (NB, CH, S)
        return y

    @staticmethod
    def backward(ctx, dy):
        x_, y, mr, gamma = ctx.saved_tensors
        NB, CH, S = ctx.dims
        sums = torch.empty((2, CH), dtype=torch.float32, device=x_.device)   # rows: dbeta, dgamma
        dx = torch.empty_like(x_)
        dy_c = _c(dy)
        call("icl_bn_relu_bwd", P(dy_c), P(x_), P(y), P(mr), P(gamma), P(sums), P(dx), c_int(NB), c_int(CH), c_ll(S), P(ops.reduce_ws(x_.device)),
             mbytes=28e-6 * x_.numel(), tag="NB%d CH%d S%d" % (NB, CH, S))
        return dx, sums[1], sums[0], None, None, None, None, None


_PENDING_NBT = []
_DEFER_NBT = [False]


def defer_batch_counters():
    """From here to flush_batch_counters(), `num_batches_tracked += 1` of the BatchNorms that run is collected instead of launched."""
    _DEFER_NBT[0] = True


def flush_batch_counters():
    """The collected counter increments as ONE multi-tensor add (the ICL heads run 18 BatchNorms per training step)."""
    _DEFER_NBT[0] = False
    if _PENDING_NBT:
        torch._foreach_add_(list(_PENDING_NBT), 1)
        del _PENDING_NBT[:]


def bn_relu(x, bn, training):
    if training and bn.num_batches_tracked is not None:
        if _DEFER_NBT[0]:
            _PENDING_NBT.append(bn.num_batches_tracked)
        else:
            bn.num_batches_tracked += 1
    return BnReluFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, training, bn.momentum, bn.eps)


class PlanarPointwiseFn(torch.autograd.Function):
    """1x1x1 Conv3d on planar [NB, CI, S...] maps: y[nb] = W @ x[nb] (+ b)  (pointwise conv :325, attn_convs1 :196)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x_ = _c(x.detach())
        NB, CI = x_.shape[0], x_.shape[1]
        S = x_.numel() // (NB * CI)
        w2 = _c(w.detach()).reshape(w.shape[0], CI)
        CO = w2.shape[0]
        y = torch.empty((NB, CO) + tuple(x_.shape[2:]), dtype=torch.float32, device=x.device)
        if CI <= 16 and CO <= 16:
            b_ = None if b is None else _c(b.detach())
            call("icl_planar_pw", P(x_), P(w2), c_int(CI), c_int(1), P(b_), P(y), c_int(NB), c_int(CI), c_int(CO), c_ll(S),
                 mbytes=4e-6 * NB * S * (CI + CO), tag="NB%d %d->%d S%d" % (NB, CI, CO, S))
        else:
            ops.sgemm(CO, S, CI, w2, CI, 1, x_, S, 1, y, S, 1, bias=None if b is None else b.detach(), bias_mode=2 if b is not None else 0,
                      batch=NB, sA=0, sB=CI * S, sC=CO * S)
        ctx.save_for_backward(x_, w2)
        ctx.has_bias, ctx.wshape = b is not None, w.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        x_, w2 = ctx.saved_tensors
        dy = _c(dy)
        NB, CI = x_.shape[0], x_.shape[1]
        CO = w2.shape[0]
        S = x_.numel() // (NB * CI)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x_)
            if CI <= 16 and CO <= 16:
                call("icl_planar_pw", P(dy), P(w2), c_int(1), c_int(CI), P(None), P(dx), c_int(NB), c_int(CO), c_int(CI), c_ll(S),
                     mbytes=4e-6 * NB * S * (CI + CO), tag="NB%d %d->%d S%d (dgrad)" % (NB, CO, CI, S))
            else:
                ops.sgemm(CI, S, CO, w2, 1, CI, dy, S, 1, dx, S, 1, batch=NB, sA=0, sB=CO * S, sC=CI * S)
        dw = torch.empty_like(w2)
        db = torch.empty((CO,), dtype=torch.float32, device=dy.device) if ctx.has_bias else None
        call("icl_planar_pw_wgrad", P(dy), P(x_), P(dw), P(db), c_int(NB), c_int(CO), c_int(CI), c_ll(S), P(ops.reduce_ws(x_.device)),
             mbytes=4e-6 * NB * S * (CO + CI), tag="NB%d %d->%d S%d" % (NB, CI, CO, S))
        return dx, dw.reshape(ctx.wshape), db


def planar_pointwise(x, w, b=None):
    return PlanarPointwiseFn.apply(x, w, b)
