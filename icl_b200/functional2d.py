"""torch.autograd.Functions of the 2D path (networks/unet_icl.py, networks/unet.py).

A batch of images [N,C,H,W] is handed to the 3D kernels as ONE sample with the images stacked along depth,
F32CL [1][N][H][W][C] (= channels-last NHWC, zero-copy):
  * Conv2d 3x3 = the 3x3x3 tensor-core / CUDA-core convolution kernels with the 2D weights in the centre depth plane (the
    two outer depth taps are zero), so the forward, data-gradient and weight-gradient kernels are reused as they are;
  * BatchNorm2d batch statistics = the per-(sample, channel) sums the conv epilogue already emits for InstanceNorm
    (one sample = the whole batch), followed by icl_normact_* with the affine parameters and the LeakyReLU slope.
Activations cross Function boundaries as fp32 channels-last tensors; every convolution packs its own split-bf16 operand
(this path is the parity case of BASELINE config 1, not the benchmarked one).
"""
import torch
import torch.nn.functional as F

from . import ops
from .ops import P, c_f, c_int, c_ll, call
from .precision import planes


def _nhwc(x):
    ops._require_cuda(x)
    if x.dtype != torch.float32:
        x = x.float()
    return x.permute(0, 2, 3, 1).contiguous()


def _nchw_view(x):
    return x.permute(0, 3, 1, 2)


def _w3d(w):
    """[Cout,Cin,3,3] -> [Cout,Cin,3,3,3] with the 2D taps in the centre depth plane."""
    return F.pad(w.detach().unsqueeze(2), (0, 0, 0, 0, 1, 1)).contiguous()


def _conv_fwd(xs, w3, b, N, H, W, stats):
    """xs: list of NHWC fp32 tensors (virtual concat).  Returns y [N,H,W,Cout] (as a depth-N volume of one sample)."""
    cins = [x.shape[-1] for x in xs]
    cout = w3.shape[0]
    if ops.umma_ok(cins, cout):
        pks = [ops.pack_pk(x.view(1, N, H, W, c)) for x, c in zip(xs, cins)]
        y = ops.conv3d_umma(pks, cins, ops.pack_w_umma(w3, False), b, cout, 1, N, H, W, stats)
        return y.view(N, H, W, cout), pks
    if ops.stem_ok(cins, cout):
        return ops.conv3d_stem_fwd(xs[0].view(1, N, H, W, 1), w3, b, 1, N, H, W, stats).view(N, H, W, cout), None
    y = ops.conv3d_direct([x.view(1, N, H, W, c) for x, c in zip(xs, cins)], cins, ops.repack_w_f32(w3, False), b, cout, 1, N, H, W, stats)
    return y.view(N, H, W, cout), None


def _conv_bwd(xs, pks, w3, dY, dY_pk, N, H, W, need_dx, want_bias):
    """Returns (dw2d [Cout, sum(cins), 3, 3], db or None, [dx per source] or None).  dY: NHWC fp32 (may be None when dY_pk
    serves every consumer)."""
    cins = [x.shape[-1] for x in xs]
    cout = w3.shape[0]
    cin_total = sum(cins)
    db = None
    if pks is not None and dY_pk is not None and ops.wgrad_umma_ok(cins, cout) and N % 2 == 0:
        dw3 = ops.conv3d_wgrad_umma(pks, cins, dY_pk, cout, 1, N, H, W)
    elif ops.stem_ok(cins, cout):
        dw3 = ops.conv3d_stem_wgrad(xs[0].view(1, N, H, W, 1), dY.view(1, N, H, W, cout), 1, N, H, W)
    else:
        dw3, db = ops.conv3d_wgrad([x.view(1, N, H, W, c) for x, c in zip(xs, cins)], cins, dY.view(1, N, H, W, cout), cout, 1, N, H, W,
                                   want_bias=want_bias)
    dw2 = dw3[:, :, 1].contiguous()
    dxs = None
    if need_dx:
        if dY_pk is not None and ops.umma_ok([cout], cin_total, with_stats=False) and all(c % 16 == 0 for c in cins):
            wp = ops.pack_w_umma(w3, True)
            if len(cins) == 2:
                d0, d1 = ops.conv3d_umma([dY_pk], [cout], wp, None, cin_total, 1, N, H, W, split=cins[0])
                dxs = [d0.view(N, H, W, cins[0]), d1.view(N, H, W, cins[1])]
            else:
                dxs = [ops.conv3d_umma([dY_pk], [cout], wp, None, cin_total, 1, N, H, W).view(N, H, W, cin_total)]
        else:
            dx = ops.conv3d_direct([dY.view(1, N, H, W, cout)], [cout], ops.repack_w_f32(w3, True), None, cin_total, 1, N, H, W)
            dx = dx.view(N, H, W, cin_total)
            dxs = [dx] if len(cins) == 1 else [dx[..., :cins[0]].contiguous(), dx[..., cins[0]:].contiguous()]
    return dw2, db, dxs


class Conv2dBnActFn(torch.autograd.Function):
    """Conv2d(3x3, pad 1, bias) -> BatchNorm2d -> LeakyReLU(slope) on the channel-concat of x0 [, x1]
    (ConvBlock halves, networks/unet_icl.py:46-54; the concat of UpBlock :95 is never materialised)."""

    @staticmethod
    def forward(ctx, x0, x1, w, b, gamma, beta, run_mean, run_var, training, momentum, eps, slope):
        xs = [_nhwc(x0.detach())] + ([_nhwc(x1.detach())] if x1 is not None else [])
        N, H, W, _ = xs[0].shape
        cout = w.shape[0]
        w3 = _w3d(w)
        stats = ops.zeros((1, cout, 2), torch.float64, w.device)
        y, pks = _conv_fwd(xs, w3, b.detach(), N, H, W, stats)
        S = N * H * W
        if training:
            mr = ops.instnorm_finalize(stats, 1, cout, S, eps)
            if run_mean is not None:  # running statistics: momentum update with the UNBIASED batch variance (torch semantics)
                mean = (stats[0, :, 0] / S)
                var_b = (stats[0, :, 1] / S - mean * mean).clamp_min(0.0)
                run_mean.mul_(1.0 - momentum).add_(mean.float(), alpha=momentum)
                run_var.mul_(1.0 - momentum).add_((var_b * (S / max(S - 1, 1))).float(), alpha=momentum)
        else:
            mr = torch.stack([run_mean, torch.rsqrt(run_var + eps)], dim=1).reshape(1, cout, 2).contiguous()
        a = torch.empty_like(y)
        call("icl_normact_fwd", P(y), P(mr), P(gamma.detach()), P(beta.detach()), c_f(slope), P(a), P(None), c_int(0), c_int(1), c_int(cout),
             c_ll(S))
        ctx.save_for_backward(*xs, w3, y, mr, gamma.detach(), beta.detach())
        ctx.n_src, ctx.pks, ctx.dims, ctx.slope, ctx.training = len(xs), pks, (N, H, W), slope, training
        return _nchw_view(a)

    @staticmethod
    def backward(ctx, dA):
        saved = ctx.saved_tensors
        xs, (w3, y, mr, gamma, beta) = list(saved[:ctx.n_src]), saved[ctx.n_src:]
        N, H, W = ctx.dims
        cout = w3.shape[0]
        if not ctx.training:
            raise RuntimeError("icl_b200 Conv2dBnActFn: backward through eval-mode BatchNorm is not on the reference's path")
        dA_ = _nhwc(dA)
        S = N * H * W
        red = ops.zeros((1, cout, 2), torch.float64, y.device)
        want_pk = cout % 16 == 0
        dY = torch.empty_like(y)
        pk = ops.empty_pk(1, cout, N, H, W, y.device) if want_pk else None
        db = ops.zeros((cout,), torch.float32, y.device) if cout % 8 == 0 else None
        call("icl_normact_bwd", P(dA_), P(y), P(mr), P(gamma), P(beta), c_f(ctx.slope), P(red), P(dY), P(pk), c_int(1 if planes() == 2 else 0),
             P(db), c_int(1), c_int(cout), c_ll(S))
        dgamma, dbeta = red[0, :, 1].float(), red[0, :, 0].float()
        need_dx = [ctx.needs_input_grad[0], ctx.n_src == 2 and ctx.needs_input_grad[1]]
        dw, db2, dxs = _conv_bwd(xs, ctx.pks, w3, dY, pk, N, H, W, any(need_dx), want_bias=db is None)
        if db is None:
            db = db2
        dx0 = _nchw_view(dxs[0]) if need_dx[0] else None
        dx1 = _nchw_view(dxs[1]) if need_dx[1] else None
        return dx0, dx1, dw, db, dgamma, dbeta, None, None, None, None, None, None


def conv_bn_act(x0, x1, conv, bn, slope=0.01):
    if bn.training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1
    return Conv2dBnActFn.apply(x0, x1, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.training, bn.momentum,
                               bn.eps, slope)


class Conv2dFn(torch.autograd.Function):
    """Plain Conv2d(3x3, pad 1, bias): `out_conv` 16 -> K (networks/unet_icl.py:177-178)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x_ = _nhwc(x.detach())
        N, H, W, _ = x_.shape
        w3 = _w3d(w)
        y, pks = _conv_fwd([x_], w3, b.detach(), N, H, W, None)
        ctx.save_for_backward(x_, w3)
        ctx.pks, ctx.dims = pks, (N, H, W)
        return _nchw_view(y)

    @staticmethod
    def backward(ctx, dy):
        x_, w3 = ctx.saved_tensors
        N, H, W = ctx.dims
        cout = w3.shape[0]
        dY = _nhwc(dy)
        dY_pk = ops.pack_pk(dY.view(1, N, H, W, cout)) if cout % 16 == 0 else None
        dw, db, dxs = _conv_bwd([x_], ctx.pks, w3, dY, dY_pk, N, H, W, ctx.needs_input_grad[0], want_bias=True)
        if db is None:
            db = torch.empty((cout,), dtype=torch.float32, device=dY.device)
            call("icl_colsum", P(dY), P(db), c_ll(N * H * W), c_int(cout), c_int(0))
        return (_nchw_view(dxs[0]) if dxs is not None else None), dw, db


def conv2d_3x3(x, conv):
    return Conv2dFn.apply(x, conv.weight, conv.bias)


class MaxPool2dFn(torch.autograd.Function):
    """nn.MaxPool2d(2) (networks/unet_icl.py:66); first maximum in scan order wins ties."""

    @staticmethod
    def forward(ctx, x):
        x_ = _nhwc(x.detach())
        N, H, W, C = x_.shape
        out = torch.empty((N, H // 2, W // 2, C), dtype=torch.float32, device=x_.device)
        idx = torch.empty((N, H // 2, W // 2, C), dtype=torch.uint8, device=x_.device)
        call("icl_maxpool2d_fwd", P(x_), P(out), P(idx), P(None), c_int(0), c_int(N), c_int(C), c_int(H), c_int(W))
        ctx.save_for_backward(idx)
        ctx.dims = (N, H, W, C)
        return _nchw_view(out)

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        N, H, W, C = ctx.dims
        dx = torch.empty((N, H, W, C), dtype=torch.float32, device=idx.device)
        call("icl_maxpool2d_bwd", P(_nhwc(dout)), P(idx), P(dx), c_int(0), c_int(N), c_int(C), c_int(H), c_int(W))
        return _nchw_view(dx)


def max_pool2d(x):
    return MaxPool2dFn.apply(x)


class UpsampleAc2dFn(torch.autograd.Function):
    """nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) (networks/unet_icl.py:84-85)."""

    @staticmethod
    def forward(ctx, x):
        x_ = _nhwc(x.detach())
        N, h, w, C = x_.shape
        out = torch.empty((N, 2 * h, 2 * w, C), dtype=torch.float32, device=x_.device)
        call("icl_upsample2x_ac2d_fwd", P(x_), P(out), P(None), c_int(0), c_int(N), c_int(C), c_int(h), c_int(w))
        ctx.dims = (N, h, w, C)
        return _nchw_view(out)

    @staticmethod
    def backward(ctx, dout):
        N, h, w, C = ctx.dims
        g = _nhwc(dout)
        dx = torch.empty((N, h, w, C), dtype=torch.float32, device=g.device)
        call("icl_upsample2x_ac2d_bwd", P(g), c_int(C), c_int(0), P(dx), c_int(0), c_int(N), c_int(C), c_int(h), c_int(w))
        return _nchw_view(dx)


def upsample2x_ac(x):
    return UpsampleAc2dFn.apply(x)


class DropoutFn(torch.autograd.Function):
    """nn.Dropout(p) on a conv output: keep-mask from an explicit byte mask (parity tests; NHWC memory order) or Philox keyed
    by a device-side seed, regenerated in backward."""

    @staticmethod
    def forward(ctx, x, p, mask, seed):
        out = ops.dropout(_nhwc(x.detach()), p, mask, seed)
        ctx.p, ctx.mask, ctx.seed = p, mask, seed
        return _nchw_view(out)

    @staticmethod
    def backward(ctx, g):
        return _nchw_view(ops.dropout(_nhwc(g), ctx.p, ctx.mask, ctx.seed)), None, None, None


def dropout(x, p, training, mask=None):
    if not training or p == 0.0:
        return x
    seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device=x.device) if mask is None else 0
    return DropoutFn.apply(x, p, mask, seed)


def conv1x1(x, conv):
    """nn.Conv2d(kernel_size=1) (UpBlock.conv1x1, networks/unet_icl.py:83,93) = Linear over the channels of NHWC rows."""
    from . import functional as Fn
    N, C, H, W = x.shape
    y = Fn.linear(x.permute(0, 2, 3, 1).reshape(N, H * W, C), conv.weight.reshape(conv.weight.shape[0], C), conv.bias)
    return y.reshape(N, H, W, -1).permute(0, 3, 1, 2)
