"""Sliding-window inference — drop-in for test_single_case (reference test_3D_BraTS.py:79-142 ≡
val_3D.test_single_case_base :15-82) and the Dice half of calculate_metric_percase (:175-187).

Same window grid (last window clamped to the border), same x->y->z accumulation order, but the volume, the
score map, the visit counts, the argmax and the Dice counts all stay on the device: one H2D of the volume and
one D2H of the label map per case instead of one of each per window.  `rank`/`world_size` shard windows
round-robin across processes; the caller sums score/cnt with an all-reduce.
"""
import math

import numpy as np
import torch

from . import ops
from .ops import P, c_int, c_ll, call


def window_starts(size, patch, stride):
    n = math.ceil((size - patch) / stride) + 1
    return [min(stride * i, size - patch) for i in range(n)]


def _accumulate(logits_ndhwc, score, cnt, xs, ys, zs):
    """logits_ndhwc: [pw, ph, pd, K] (one window, channels-last)."""
    pw, ph, pd, K = logits_ndhwc.shape
    _, W, H, D = score.shape
    call("icl_sw_accumulate", P(logits_ndhwc), c_int(K), c_int(pw), c_int(ph), c_int(pd), P(score), P(cnt), c_int(W), c_int(H), c_int(D),
         c_int(xs), c_int(ys), c_int(zs))


def _finalize(score, cnt):
    K = score.shape[0]
    S = cnt.numel()
    label = torch.empty(cnt.shape, dtype=torch.int64, device=score.device)
    call("icl_sw_finalize", P(score), P(cnt), c_int(K), c_ll(S), P(label))
    return label


def _check_device(dev):
    if dev.type != "cuda":
        raise RuntimeError("icl_b200 inference runs on CUDA only (no CPU fallback)")


def sliding_window_scores(net, image, stride_xy, stride_z, patch_size, num_classes, rank=0, world_size=1, inference_kw=False, window_batch=1):
    """Returns (score [K,ww,hh,dd], cnt [ww,hh,dd], pads) for this rank's share of the windows.  `image`: numpy / CPU tensor
    [w,h,d] (copied to the device once) or a tensor already on the device.  `window_batch` windows go through the network per
    call (the reference runs one, test_3D_BraTS.py:119-133); the per-voxel accumulation order (x, then y, then z) is unchanged."""
    dev = next(net.parameters()).device
    _check_device(dev)
    img = image if isinstance(image, torch.Tensor) else torch.as_tensor(np.asarray(image), dtype=torch.float32)
    img = img.to(device=dev, dtype=torch.float32, non_blocking=True)
    w, h, d = img.shape
    pads = []
    for s, p in zip((w, h, d), patch_size):
        tot = max(p - s, 0)
        pads.append((tot // 2, tot - tot // 2))
    if any(a + b > 0 for a, b in pads):
        img = torch.nn.functional.pad(img, (pads[2][0], pads[2][1], pads[1][0], pads[1][1], pads[0][0], pads[0][1]))
    ww, hh, dd = img.shape
    px, py, pz = patch_size
    score = torch.zeros((num_classes, ww, hh, dd), dtype=torch.float32, device=dev)
    cnt = torch.zeros((ww, hh, dd), dtype=torch.float32, device=dev)
    wins = [(xs, ys, zs) for xs in window_starts(ww, px, stride_xy) for ys in window_starts(hh, py, stride_xy)
            for zs in window_starts(dd, pz, stride_z)]
    mine = [wn for n, wn in enumerate(wins) if (n % world_size) == rank]
    nb = max(1, int(window_batch))
    with torch.no_grad():
        for i in range(0, len(mine), nb):
            grp = mine[i:i + nb]
            patch = torch.stack([img[xs:xs + px, ys:ys + py, zs:zs + pz] for xs, ys, zs in grp])[:, None]
            y1 = net(patch, inference=True) if inference_kw else net(patch)
            y1 = ops.to_ndhwc(y1)
            for j, (xs, ys, zs) in enumerate(grp):
                _accumulate(y1[j], score, cnt, xs, ys, zs)
    return score, cnt, pads


_PINNED = {}


def _label_to_host(label):
    """Device int64 label map -> numpy int64.  Class ids fit one byte, so 1 byte per voxel crosses PCIe (through a cached pinned
    buffer) instead of 8, and the widening back to the reference's int64 happens on the host."""
    if not label.is_cuda:   # host-logic tests drive the window sharding with CPU stand-in networks: nothing to transfer
        return label.numpy().astype(np.int64)
    small = label.to(torch.uint8)
    key = (small.numel(), label.device.index)
    buf = _PINNED.get(key)
    if buf is None:
        if len(_PINNED) > 4:
            _PINNED.clear()
        buf = _PINNED[key] = torch.empty((small.numel(),), dtype=torch.uint8).pin_memory()
    buf.copy_(small.reshape(-1), non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return buf.numpy().reshape(tuple(label.shape)).astype(np.int64)


def test_single_case(net, image, stride_xy, stride_z, patch_size, num_classes=1, inference_kw=False, window_batch=1):
    """numpy [w,h,d] in -> numpy int64 label map out (host-blocking, like the reference)."""
    w, h, d = image.shape
    score, cnt, pads = sliding_window_scores(net, image, stride_xy, stride_z, patch_size, num_classes, inference_kw=inference_kw,
                                             window_batch=window_batch)
    label = _finalize(score, cnt)
    label = label[pads[0][0]:pads[0][0] + w, pads[1][0]:pads[1][0] + h, pads[2][0]:pads[2][0] + d]
    return _label_to_host(label)


test_single_case.__test__ = False  # not a pytest test


def test_single_case_sharded(net, image, stride_xy, stride_z, patch_size, num_classes=1, group=None, inference_kw=False, window_batch=1):
    """Window-sharded test_single_case: every rank of the initialised process group evaluates its round-robin
    share of the windows, the score map and the visit counts are summed with one all-reduce each (the only exchange step of the
    path), and every rank finalises the same label map.  The sums are formed in a different order than the single-process
    accumulation, so scores agree to fp32 rounding and the argmax can differ only where two class scores tie to ~1e-7.
    Without an initialised process group this is test_single_case."""
    import torch.distributed as dist
    on = dist.is_available() and dist.is_initialized()
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if on else (0, 1)
    w, h, d = image.shape
    score, cnt, pads = sliding_window_scores(net, image, stride_xy, stride_z, patch_size, num_classes, rank=rank, world_size=world,
                                             inference_kw=inference_kw, window_batch=window_batch)
    if world > 1:
        dist.all_reduce(score, group=group)
        dist.all_reduce(cnt, group=group)
    label = _finalize(score, cnt)
    label = label[pads[0][0]:pads[0][0] + w, pads[1][0]:pads[1][0] + h, pads[2][0]:pads[2][0] + d]
    return _label_to_host(label)


test_single_case_sharded.__test__ = False  # not a pytest test


def dice_metric(pred, gt):
    """Dice of (pred>0) vs (gt>0) with exact integer counts; the empty-set conventions of
    calculate_metric_percase (test_3D_BraTS.py:175-187).  Returns (dice, (|A∩B|, |A|, |B|))."""
    p = torch.as_tensor(pred).long().contiguous()
    g = torch.as_tensor(gt).long().contiguous()
    if not (p.is_cuda and g.is_cuda):
        raise RuntimeError("icl_b200.inference.dice_metric runs on CUDA tensors only")
    counts = torch.zeros((3,), dtype=torch.int64, device=p.device)
    call("icl_dice_counts", P(p), P(g), c_ll(p.numel()), P(counts))
    inter, na, nb = (int(v) for v in counts.cpu())
    if na > 0 and nb > 0:
        dice = 2.0 * inter / float(na + nb)
    elif na == 0 and nb == 0:
        dice = 1.0
    else:
        dice = 0.0
    return dice, (inter, na, nb)


def cal_metric_dice(gt, pred):
    """Dice half of cal_metric (val_3D.py:85-97) for one class: binarise, exact integer counts on the device, the reference's
    empty-set conventions (both empty -> 1, one empty -> 0).  HD95 is out of scope (MedPy / scipy CPU code)."""
    return dice_metric(pred, gt)[0]


def test_all_case(net, cases, num_classes=2, patch_size=(96, 96, 96), stride_xy=64, stride_z=64, window_batch=4, inference_kw=False,
                  group=None):
    """In-training validation driver — test_all_case_base (val_3D.py:100-118; called every 200 iterations by
    train_inherent_consistent_unet_3D_BraTS.py:135-137) on in-memory cases instead of an h5 list: `cases` yields (image [w,h,d],
    label [w,h,d]) arrays or tensors.  Every case runs the device-resident sliding window (windows batched, sharded over the ranks
    of an initialised process group), argmax and the per-class Dice counts stay on the GPU; nothing but one float per (case, class)
    crosses PCIe.  Returns metric_cal like the reference: a list over classes 1..num_classes-1 of per-case (dice, hd95) tuples, with
    hd95 = nan (out of scope)."""
    import torch.distributed as dist
    on = dist.is_available() and dist.is_initialized()
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if on else (0, 1)
    dev = next(net.parameters()).device
    metric_cal = [[] for _ in range(num_classes - 1)]
    for image, label in cases:
        w, h, d = image.shape
        score, cnt, pads = sliding_window_scores(net, image, stride_xy, stride_z, patch_size, num_classes, rank=rank, world_size=world,
                                                 inference_kw=inference_kw, window_batch=window_batch)
        if world > 1:
            dist.all_reduce(score, group=group)
            dist.all_reduce(cnt, group=group)
        pred = _finalize(score, cnt)[pads[0][0]:pads[0][0] + w, pads[1][0]:pads[1][0] + h, pads[2][0]:pads[2][0] + d]
        gt = (label if isinstance(label, torch.Tensor) else torch.as_tensor(np.asarray(label))).to(dev).long()
        for i in range(1, num_classes):
            metric_cal[i - 1].append((cal_metric_dice((gt == i).long(), (pred == i).long()), float("nan")))
    return metric_cal


test_all_case.__test__ = False  # not a pytest test
