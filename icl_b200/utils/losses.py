"""Loss callables of the ICL training loops — drop-in for the reference's utils/losses.py (the eight names
on the hot path) plus CrossEntropyLoss (torch.nn.modules.loss in the reference).

All return 0-dim differentiable CUDA tensors.  Under the hood one streaming "class statistics" kernel family
(icl_b200/csrc/losses.cu) fuses softmax, CE, the Dice partial sums and — for the multi-scale losses — the
trilinear interpolation of the coarse class map to the label grid, so no [B,K,96^3] temporary is written.
No .item() host syncs (the reference's DiceLoss does K of them per call, utils/losses.py:229).
"""
import torch
import torch.nn as nn

from .. import lanes, ops
from ..ops import P, c_d, c_f, c_int, c_ll, call


def _fan(K, device, thunks, tensors):
    """The scales of a multi-scale loss are independent: with many classes (K >= 8: every scale streams >100 MB at 96^3 and one
    kernel alone reaches a third of the HBM bandwidth) each runs on a lane of its own (icl_b200/lanes.py), summed in the reference's
    order after the join.  With few classes the kernels are 25-35 us and the fork / join costs what it saves (measured, config 2)."""
    if K < 8:
        return [t() for t in thunks]
    return lanes.fan(device, "loss", thunks, tensors)


def _layout(src):
    """Returns (tensor, planar flag): channels-last 5-D tensors are consumed in place, anything else as planar.
    4-D [B,K,H,W] inputs (the 2D path) are handled as depth-1 volumes."""
    if src.dim() == 4:
        src = src.unsqueeze(2)
    if src.dim() != 5:
        raise RuntimeError("icl_b200 losses expect [B,K,D,H,W] or [B,K,H,W] inputs")
    if src.dtype != torch.float32:
        src = src.float()
    if src.shape[1] > 1 and src.permute(0, 2, 3, 4, 1).is_contiguous():
        return src, 0
    return src.contiguous(), 1


class _ClassStatsFn(torch.autograd.Function):
    """(src [B,K,r,r,r], labels | soft-target logits) -> (ce, dice) scalars on the grid `size`."""

    @staticmethod
    def forward(ctx, src, labels, tgt, size, is_prob, class_w):
        ops._require_cuda(src)
        s, planar = _layout(src.detach())
        B, K = s.shape[0], s.shape[1]
        rz, ry, rx = s.shape[2:]
        Z, Y, X = size if len(size) == 3 else (1,) + tuple(size)
        if K > 16:
            raise RuntimeError("icl_b200 losses support up to 16 classes (got %d)" % K)
        if labels is not None:
            labels = labels.detach().long().contiguous()
            if labels.numel() != B * Z * Y * X:
                raise RuntimeError("labels shape %s does not match loss grid %s" % (tuple(labels.shape), (B, Z, Y, X)))
        if tgt is not None:
            tgt = ops.to_ndhwc(tgt.detach() if tgt.dim() == 5 else tgt.detach().unsqueeze(2))
        sums = ops.zeros((3 * K + 1,), torch.float64, s.device)
        out2 = torch.empty((2,), dtype=torch.float32, device=s.device)
        call("icl_class_stats_fwd", P(s), c_int(planar), c_int(rz), c_int(ry), c_int(rx), c_int(B), c_int(K), c_int(Z), c_int(Y), c_int(X),
             P(labels), P(tgt), c_int(1 if is_prob else 0), P(class_w), P(sums), P(out2),
             mbytes=1e-6 * (s.numel() * 4 + B * Z * Y * X * (8 if labels is not None else 4 * K)), tag="K%d r%d->%d %s" % (K, rx, X, "labels" if labels is not None else "soft"))
        ctx.save_for_backward(s, labels, tgt, sums, class_w)
        ctx.meta = (planar, (rz, ry, rx), B, K, (Z, Y, X), is_prob, src.shape)
        ctx.set_materialize_grads(False)
        return out2[0], out2[1]

    @staticmethod
    def backward(ctx, g_ce, g_dice):
        s, labels, tgt, sums, class_w = ctx.saved_tensors
        planar, (rz, ry, rx), B, K, (Z, Y, X), is_prob, shape = ctx.meta
        if g_ce is None and g_dice is None:
            return None, None, None, None, None, None
        direct = (rz, ry, rx) == (Z, Y, X)
        ds = torch.empty_like(s) if direct else torch.zeros_like(s)
        ws = None if direct else torch.empty(B * K * Z * Y * X, dtype=torch.float32, device=s.device)
        gc = None if g_ce is None else g_ce.detach().float().contiguous()
        gd = None if g_dice is None else g_dice.detach().float().contiguous()
        call("icl_class_stats_bwd", P(s), c_int(planar), c_int(rz), c_int(ry), c_int(rx), c_int(B), c_int(K), c_int(Z), c_int(Y), c_int(X),
             P(labels), P(tgt), c_int(1 if is_prob else 0), P(class_w), P(sums), P(gc), P(gd), c_f(1.0), c_f(1.0), P(ds), P(ws),
             mbytes=1e-6 * (2 * s.numel() * 4 + B * Z * Y * X * (8 if labels is not None else 4 * K)), tag="K%d r%d->%d %s" % (K, rx, X, "labels" if labels is not None else "soft"))
        if len(shape) == 4:
            ds = ds.squeeze(2)
        return ds, None, None, None, None, None


def seg_ce_dice(logits, labels, size=None, class_w=None):
    """Fused nn.CrossEntropyLoss()(logits, labels) and DiceLoss(K)(softmax(logits), labels) -> (ce, dice).
    `logits` may be coarser than `size`; it is then trilinearly interpolated (align_corners=False) on the fly."""
    size = tuple(size) if size is not None else tuple(logits.shape[2:])
    return _ClassStatsFn.apply(logits, labels, None, size, False, class_w)


def soft_dice(input_logits, target_logits, size=None):
    """softmax_dice_loss (utils/losses.py:42-59) with optional fused interpolation of the input."""
    size = tuple(size) if size is not None else tuple(target_logits.shape[2:])
    return _ClassStatsFn.apply(input_logits, None, target_logits, size, False, None)[1]


class CrossEntropyLoss(nn.Module):
    """nn.CrossEntropyLoss() with default arguments (mean over voxels), train_..._BraTS.py:87,107."""

    def forward(self, logits, target):
        return seg_ce_dice(logits, target)[0]


class DiceLoss(nn.Module):
    """utils/losses.py:195-231.  inputs: probabilities (softmax=False) or logits (softmax=True); target [B,1,...]."""

    def __init__(self, n_classes):
        super().__init__()
        self.n_classes = n_classes

    def forward(self, inputs, target, weight=None, softmax=False):
        if inputs.shape[1] != self.n_classes:
            raise AssertionError("predict %s & n_classes %d do not match" % (tuple(inputs.shape), self.n_classes))
        if target.dim() == inputs.dim():
            assert target.shape[1] == 1, "target must be [B,1,...]"
            target = target[:, 0]
        assert tuple(inputs.shape[2:]) == tuple(target.shape[1:]) and inputs.shape[0] == target.shape[0], \
            "predict {} & target {} shape do not match".format(tuple(inputs.shape), tuple(target.shape))
        cw = None
        if weight is not None:
            cw = torch.as_tensor(weight, dtype=torch.float32, device=inputs.device).contiguous()
        return _ClassStatsFn.apply(inputs, target, None, tuple(inputs.shape[2:]), not softmax, cw)[1]


class AuxLoss3D(nn.Module):
    """utils/losses.py:254-271: mean over scales of CE + Dice on feature maps interpolated to 96^3."""

    def __init__(self, n_classes, resize=(96, 96, 96)):
        super().__init__()
        self.n_classes = n_classes
        self.resize = tuple(resize)

    def forward(self, feat_maps, labels):
        ce_sum, dice_sum = None, None
        res = _fan(feat_maps[0].shape[1], labels.device, [lambda fm=fm: seg_ce_dice(fm, labels, self.resize) for fm in feat_maps],
                   list(feat_maps) + [labels])
        for ce, dc in res:
            ce_sum = ce if ce_sum is None else ce_sum + ce
            dice_sum = dc if dice_sum is None else dice_sum + dc
        n = len(feat_maps)
        return ce_sum / n + dice_sum / n


class PseudoSoftLoss3D(nn.Module):
    """utils/losses.py:287-299: softmax-Dice of interpolated feature maps against the detached prediction."""

    def __init__(self, n_classes=None, resize=(96, 96, 96)):
        super().__init__()
        self.resize = tuple(resize)

    def forward(self, feat_maps, predicts):
        tgt = predicts.detach()
        tot = None
        for d in _fan(tgt.shape[1], tgt.device, [lambda fm=fm: soft_dice(fm, tgt, self.resize) for fm in feat_maps], list(feat_maps) + [tgt]):
            tot = d if tot is None else tot + d
        return tot / len(feat_maps)


class AuxLoss(AuxLoss3D):
    """utils/losses.py:233-251 (2D): feature maps bilinearly resized (align_corners=False) to `resize`, CE + Dice(softmax=True)."""

    def __init__(self, n_classes, resize=(224, 224)):
        super().__init__(n_classes, tuple(resize))


class PseudoSoftLoss(PseudoSoftLoss3D):
    """utils/losses.py:273-285 (2D)."""

    def __init__(self, n_classes=None, resize=(224, 224)):
        super().__init__(n_classes, tuple(resize))


class _SoftmaxMseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ops._require_cuda(a)
        a_ = a.detach().float().contiguous()
        b_ = b.detach().float().contiguous()
        B, K = a_.shape[0], a_.shape[1]
        S = a_.numel() // (B * K)
        if K > 16:
            raise RuntimeError("icl_b200 softmax_mse supports up to 16 classes")
        acc = ops.zeros((1,), torch.float64, a.device)
        call("icl_softmax_mse", P(a_), P(b_), c_int(B), c_int(K), c_ll(S), P(acc), P(None), c_f(1.0), P(None))
        out = torch.empty((1,), dtype=torch.float32, device=a.device)
        call("icl_scale_to_float", P(acc), c_d(1.0 / float(B * K * S)), P(out))
        ctx.save_for_backward(a_, b_)
        ctx.dims = (B, K, S)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        a_, b_ = ctx.saved_tensors
        B, K, S = ctx.dims
        da = torch.empty_like(a_)
        call("icl_softmax_mse", P(a_), P(b_), c_int(B), c_int(K), c_ll(S), P(None), P(g.detach().float().contiguous()), c_f(1.0), P(da))
        return da, None


def softmax_mse_loss(input_logits, target_logits, sigmoid=False):
    """utils/losses.py:68-90 (sigmoid=False branch): mean over scales of MSE between softmaxes; targets detached."""
    if sigmoid:
        raise NotImplementedError("icl_b200 softmax_mse_loss implements the sigmoid=False branch the ICL loops use")
    tot = None
    pairs = list(zip(input_logits, target_logits))
    dev = pairs[0][0].device if pairs else torch.device("cpu")
    for m in _fan(pairs[0][0].shape[1] if pairs else 0, dev, [lambda a=a, b=b: _SoftmaxMseFn.apply(a, b.detach()) for a, b in pairs],
                  [t for ab in pairs for t in ab]):
        tot = m if tot is None else tot + m
    return tot / len(input_logits)


def softmax_dice_loss(input_logits, target_logits):
    """utils/losses.py:42-59."""
    assert input_logits.size() == target_logits.size()
    return soft_dice(input_logits, target_logits)
