"""CPU restatement of the ICL hot path (TEST INFRASTRUCTURE — see oracle/__init__.py).

Functional form over a flat ``state_dict``-style mapping ``P`` (name -> tensor) instead of
the reference's nn.Module tree; every function cites the reference lines it follows
(paths relative to /root/reference/code).  Arithmetic is delegated to torch CPU ops — the
same third-party dependency that holds all of the reference's arithmetic.  Pinned against
the live reference by tests/test_oracle_vs_reference.py and against the committed fixtures
in tests/golden/ by tests/test_oracle_golden.py.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# randomness sources (Dropout / DropPath).  SURVEY.md §7.3(5): masks are inputs to parity.
# --------------------------------------------------------------------------------------


class NoRand:
    """eval(): Dropout and DropPath are the identity."""

    def dropout(self, x, p):
        return x

    def droppath(self, x, p):
        return x


class TorchRand:
    """train(): draws from the global torch generator in the reference's order
    (unet_3D_icl.py:110-147; MONAI DropPath: one Bernoulli(keep) per sample / keep).
    Records every mask drawn so the same masks can be replayed on another implementation."""

    def __init__(self):
        self.record = []

    def dropout(self, x, p):
        noise = torch.empty_like(x).bernoulli_(1.0 - p).div_(1.0 - p)
        self.record.append(("dropout", noise))
        return x * noise

    def droppath(self, x, p):
        if p == 0.0:
            return x
        keep = 1.0 - p
        r = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep).div_(keep)
        self.record.append(("droppath", r))
        return x * r


class ReplayRand:
    """Replays masks recorded by TorchRand (same call order)."""

    def __init__(self, record):
        self.record = list(record)
        self.pos = 0

    def _next(self, kind):
        k, m = self.record[self.pos]
        assert k == kind, (k, kind)
        self.pos += 1
        return m

    def dropout(self, x, p):
        return x * self._next("dropout").to(x.dtype)

    def droppath(self, x, p):
        if p == 0.0:
            return x
        return x * self._next("droppath").to(x.dtype)


# --------------------------------------------------------------------------------------
# 3D U-Net backbone  (networks/unet_3D.py:71-94, networks/unet_3D_icl.py:100-117)
# --------------------------------------------------------------------------------------


def conv_in_relu(P, prefix, x):
    """One half of UnetConv3: Conv3d 3^3 s1 p1 + bias -> InstanceNorm3d(affine=False, eps 1e-5,
    biased var) -> ReLU.  networks/utils.py:104-109."""
    y = F.conv3d(x, P[prefix + ".0.weight"], P[prefix + ".0.bias"], stride=1, padding=1)
    y = F.instance_norm(y, eps=1e-5)
    return F.relu(y)


def unetconv3(P, prefix, x):
    """UnetConv3.forward, networks/utils.py:120-123."""
    return conv_in_relu(P, prefix + ".conv2", conv_in_relu(P, prefix + ".conv1", x))


def unetup3_ct(P, prefix, skip, coarse):
    """UnetUp3_CT.forward, networks/utils.py:271-276: trilinear x2 (align_corners=False), zero
    pad of the skip by offset//2 (0 for even sizes), cat([skip, up], 1), UnetConv3."""
    up = F.interpolate(coarse, scale_factor=(2, 2, 2), mode="trilinear", align_corners=False)
    offset = up.shape[2] - skip.shape[2]
    pad = 2 * [offset // 2, offset // 2, 0]
    skip = F.pad(skip, pad)
    return unetconv3(P, prefix + ".conv", torch.cat([skip, up], 1))


def backbone3d(P, x, rand=None, p_drop=0.3):
    """Shared body of unet_3D.forward (unet_3D.py:71-94) and each branch of
    unet_3D_icl.forward (unet_3D_icl.py:100-117 / 122-139).
    Returns (final logits, center after dropout1, up4, up3)."""
    rand = rand or NoRand()
    c1 = unetconv3(P, "conv1", x)
    c2 = unetconv3(P, "conv2", F.max_pool3d(c1, 2))
    c3 = unetconv3(P, "conv3", F.max_pool3d(c2, 2))
    c4 = unetconv3(P, "conv4", F.max_pool3d(c3, 2))
    center = unetconv3(P, "center", F.max_pool3d(c4, 2))
    center = rand.dropout(center, p_drop)
    up4 = unetup3_ct(P, "up_concat4", c4, center)
    up3 = unetup3_ct(P, "up_concat3", c3, up4)
    up2 = unetup3_ct(P, "up_concat2", c2, up3)
    up1 = unetup3_ct(P, "up_concat1", c1, up2)
    up1 = rand.dropout(up1, p_drop)
    final = F.conv3d(up1, P["final.weight"], P["final.bias"])
    return final, center, up4, up3


def unet_3d_forward(P, x, rand=None):
    """unet_3D.forward (unet_3D.py:71-94)."""
    return backbone3d(P, x, rand)[0]


# --------------------------------------------------------------------------------------
# ICL heads (networks/unet_3D_icl.py:155-345)
# --------------------------------------------------------------------------------------

ICL_HEADS_3D = (16, 8, 4)  # unet_3D_icl.py:86,95
ICL_DROP_PATH = 0.02  # dpr[1] of linspace(0, 0.1, 6), unet_3D_icl.py:175,193


def _ln(x, P, prefix):
    return F.layer_norm(x, (x.shape[-1],), P[prefix + ".weight"], P[prefix + ".bias"], 1e-5)


def _mlp(x, P, prefix):
    """MLP.forward (unet_3D_icl.py:308-314): fc1 -> GELU(erf) -> fc2 (Dropout p=0)."""
    h = F.gelu(F.linear(x, P[prefix + ".fc1.weight"], P[prefix + ".fc1.bias"]))
    return F.linear(h, P[prefix + ".fc2.weight"], P[prefix + ".fc2.bias"])


def query_attention(P, prefix, q, x, num_heads):
    """Query_Attention.forward (unet_3D_icl.py:283-297) in explicit index form.

    fc_q(q) [B,K,C] is *flat-reinterpreted* as [B,H,K,hd] (no head transpose, :287); k and v are
    the usual per-head splits of fc_kv (:288-289); the returned map is the scaled logits BEFORE
    softmax, laid out [B,K,H,N] (:296); (softmax·v) [B,H,K,hd] is flat-reinterpreted as [B,K,C]
    before proj (:293)."""
    B, N, C = x.shape
    K = q.shape[1]
    hd = C // num_heads
    scale = hd ** -0.5
    ql = F.linear(q, P[prefix + ".fc_q.weight"], P[prefix + ".fc_q.bias"])
    qh = ql.reshape(B, K * C).reshape(B, num_heads, K, hd)
    kv = F.linear(x, P[prefix + ".fc_kv.weight"], P[prefix + ".fc_kv.bias"])
    k = kv[..., :C].reshape(B, N, num_heads, hd)
    v = kv[..., C:].reshape(B, N, num_heads, hd)
    logits = torch.einsum("bhkd,bnhd->bhkn", qh, k) * scale
    prob = logits.softmax(dim=-1)
    xv = torch.einsum("bhkn,bnhd->bhkd", prob, v).reshape(B, K * C).reshape(B, K, C)
    out = F.linear(xv, P[prefix + ".proj.weight"], P[prefix + ".proj.bias"])
    return out, logits.permute(0, 2, 1, 3)


def class_decoder(P, prefix, query, feat, num_heads, rand):
    """Class_Decoder.forward (unet_3D_icl.py:260-268).  Note ``q = q + dp(q)`` doubles the
    query (not a residual with the input) and norm3/mlp2 act over the spatial axis N."""
    q, a = query_attention(P, prefix + ".attn", _ln(query, P, prefix + ".norm1_query"),
                           _ln(feat, P, prefix + ".norm1"), num_heads)
    q = q + rand.droppath(q, ICL_DROP_PATH)
    q = q + rand.droppath(_mlp(_ln(q, P, prefix + ".norm2"), P, prefix + ".mlp"), ICL_DROP_PATH)
    a = a + rand.droppath(a, ICL_DROP_PATH)
    a = a + rand.droppath(_mlp(_ln(a, P, prefix + ".norm3"), P, prefix + ".mlp2"), ICL_DROP_PATH)
    return q, a


def _bn_train(x, P, prefix, training, momentum=0.1, eps=1e-5):
    return F.batch_norm(x, P.get(prefix + ".running_mean"), P.get(prefix + ".running_var"),
                        P[prefix + ".weight"], P[prefix + ".bias"], training, momentum, eps)


def separable_conv3d(P, prefix, x, training):
    """SeparableConv3d(relu_first=False).forward (unet_3D_icl.py:334-345): depthwise 3^3
    (groups=C, no bias) -> BN3d -> ReLU -> pointwise 1^3 (no bias) -> BN3d -> ReLU."""
    C = x.shape[1]
    y = F.conv3d(x, P[prefix + ".block.depthwise.weight"], None, 1, 1, 1, groups=C)
    y = F.relu(_bn_train(y, P, prefix + ".block.bn_depth", training))
    y = F.conv3d(y, P[prefix + ".block.pointwise.weight"], None)
    y = F.relu(_bn_train(y, P, prefix + ".block.bn_point", training))
    return y


def inherent_consistent(P, prefix, feats, guided_Q=None, modal="labeled", heads=ICL_HEADS_3D,
                        rand=None, training=True):
    """InherentConsistent.forward (unet_3D_icl.py:202-242).  ``modal='labeled'`` starts from
    the learnable guided_Q and chains query_convs; ``'unlabeled'`` uses the labeled pass's
    batch-mean queries per level.  Spatial side comes from the feature map itself instead of
    int(np.cbrt(N)) (:215)."""
    rand = rand or NoRand()
    feat_maps, updated_Qs = [], []
    B = feats[0].shape[0]
    next_Q = P[prefix + ".guided_Q"].expand(B, -1, -1) if modal == "labeled" else None
    for i, f in enumerate(feats):
        d, h, w = f.shape[2:]
        tok = F.conv3d(f, P["%s.proj_layers.%d.weight" % (prefix, i)], P["%s.proj_layers.%d.bias" % (prefix, i)])
        tok = _ln(tok.flatten(2).transpose(1, 2), P, "%s.norm_layers.%d" % (prefix, i))
        q_in = next_Q if modal == "labeled" else guided_Q[i].expand(B, -1, -1)
        q, a = class_decoder(P, "%s.class_decoders.%d" % (prefix, i), q_in, tok, heads[i], rand)
        bs, K, H, N = a.shape
        a = a.contiguous().view(bs * K, H, d, h, w)
        a = separable_conv3d(P, "%s.attn_convs0.%d" % (prefix, i), a, training)
        fm = F.conv3d(a, P["%s.attn_convs1.%d.weight" % (prefix, i)], P["%s.attn_convs1.%d.bias" % (prefix, i)])
        feat_maps.append(fm.reshape(bs, K, d, h, w))
        wq = P["%s.query_convs.%d.weight" % (prefix, i)]
        next_Q = F.linear(q, wq[:, :, 0], P["%s.query_convs.%d.bias" % (prefix, i)])  # Conv1d k=1 (:220-221)
        updated_Qs.append(q.mean(dim=0, keepdim=True))
    return feat_maps, updated_Qs


def unet_3d_icl_forward(P, x_lab, x_unlab=None, inference=None, rand=None, training=True):
    """unet_3D_icl.forward (unet_3D_icl.py:99-148)."""
    rand = rand or NoRand()
    final_lab, center_lab, up4_lab, up3_lab = backbone3d(P, x_lab, rand)
    if inference:
        return final_lab
    final_unlab, center_unlab, up4_unlab, up3_unlab = backbone3d(P, x_unlab, rand)
    feats_lab = [center_lab, up4_lab, up3_lab]
    feats_unlab = [center_unlab, up4_unlab, up3_unlab]
    maps_lab, Qs_lab = inherent_consistent(P, "sspa", feats_lab, None, "labeled", rand=rand, training=training)
    maps_consis, _ = inherent_consistent(P, "sspa", feats_unlab, None, "labeled", rand=rand, training=training)
    maps_unlab, _ = inherent_consistent(P, "uscl", feats_unlab, Qs_lab, "unlabeled", rand=rand, training=training)
    return final_lab, final_unlab, maps_lab, maps_unlab, maps_consis


# --------------------------------------------------------------------------------------
# losses (utils/losses.py)
# --------------------------------------------------------------------------------------

SMOOTH = 1e-5


def ce_loss(logits, labels):
    """nn.CrossEntropyLoss() mean over voxels (train_..._BraTS.py:87,107)."""
    return F.cross_entropy(logits, labels.long())


def dice_loss(inputs, target, n_classes, softmax=False):
    """DiceLoss.forward (utils/losses.py:195-231): one-hot by equality, squares in the
    denominator, sums over batch+space per class, mean over classes."""
    if softmax:
        inputs = torch.softmax(inputs, dim=1)
    loss = 0.0
    for i in range(n_classes):
        t = (target[:, 0] == i).float()
        p = inputs[:, i]
        inter = torch.sum(p * t)
        loss = loss + (1 - (2 * inter + SMOOTH) / (torch.sum(p * p) + torch.sum(t * t) + SMOOTH))
    return loss / n_classes


def aux_loss_3d(feat_maps, labels, n_classes, size=(96, 96, 96)):
    """AuxLoss3D.forward (utils/losses.py:261-271)."""
    ce, dc = 0.0, 0.0
    for fm in feat_maps:
        up = F.interpolate(fm.float(), size=list(size), mode="trilinear", align_corners=False)
        ce = ce + ce_loss(up, labels)
        dc = dc + dice_loss(up, labels.unsqueeze(1), n_classes, softmax=True)
    n = len(feat_maps)
    return ce / n + dc / n


def softmax_dice(input_logits, target_logits):
    """softmax_dice_loss + dice_loss1 (utils/losses.py:42-59, 22-30): no squares."""
    p = F.softmax(input_logits, dim=1)
    q = F.softmax(target_logits, dim=1)
    K = input_logits.shape[1]
    tot = 0.0
    for i in range(K):
        inter = torch.sum(p[:, i] * q[:, i])
        tot = tot + (1 - (2 * inter + SMOOTH) / (torch.sum(p[:, i]) + torch.sum(q[:, i]) + SMOOTH))
    return tot / K


def pseudo_soft_loss_3d(feat_maps, predicts, size=(96, 96, 96)):
    """PseudoSoftLoss3D.forward (utils/losses.py:290-299); target detached."""
    tgt = predicts.detach()
    tot = 0.0
    for fm in feat_maps:
        up = F.interpolate(fm.float(), size=list(size), mode="trilinear", align_corners=False)
        tot = tot + softmax_dice(up, tgt)
    return tot / len(feat_maps)


def softmax_mse_loss(input_logits, target_logits):
    """softmax_mse_loss (utils/losses.py:68-90), sigmoid=False branch; targets detached."""
    tot = 0.0
    for a, b in zip(input_logits, target_logits):
        tot = tot + torch.mean((F.softmax(a, dim=1) - F.softmax(b.detach(), dim=1)) ** 2)
    return tot / len(input_logits)


BRATS_WEIGHTS = dict(dice=1.0, ce=1.0, aux=1.0, pse=1.0, cons=10.0)  # train_..._BraTS.py:112
AMOS_WEIGHTS = dict(dice=1.0, ce=1.0, aux=1.0, pse=0.1, cons=10.0)  # train_..._AMOS22.py:230


def icl_losses(outputs, labels_lab, n_classes, weights=BRATS_WEIGHTS, size=(96, 96, 96)):
    """The five loss terms and their weighted sum (train_..._BraTS.py:105-112)."""
    final_lab, final_unlab, maps_lab, maps_unlab, maps_consis = outputs
    soft = torch.softmax(final_lab, dim=1)
    L = OrderedDict()
    L["ce"] = ce_loss(final_lab, labels_lab)
    L["dice"] = dice_loss(soft, labels_lab.unsqueeze(1), n_classes)
    L["aux"] = aux_loss_3d(maps_lab, labels_lab, n_classes, size)
    L["pse"] = pseudo_soft_loss_3d(maps_unlab, final_unlab, size)
    L["cons"] = softmax_mse_loss(maps_unlab, maps_consis)
    L["total"] = (weights["dice"] * L["dice"] + weights["ce"] * L["ce"] + weights["aux"] * L["aux"]
                  + weights["pse"] * L["pse"] + weights["cons"] * L["cons"])
    return L


# --------------------------------------------------------------------------------------
# training step + SGD (train_..._BraTS.py:103-119; torch.optim.SGD semantics)
# --------------------------------------------------------------------------------------


def make_params(state, requires_grad=True):
    P = OrderedDict()
    for k, v in state.items():
        t = v.detach().clone()
        if requires_grad and t.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            t.requires_grad_(True)
        P[k] = t
    return P


def train_step_3d(P, volume_batch, label_batch, labeled_bs, n_classes, weights=BRATS_WEIGHTS, rand=None,
                  size=(96, 96, 96)):
    """forward + losses + backward; returns (losses, grads) — grads[name] is None exactly where
    the reference leaves .grad None."""
    outputs = unet_3d_icl_forward(P, volume_batch[:labeled_bs], volume_batch[labeled_bs:], rand=rand)
    L = icl_losses(outputs, label_batch[:labeled_bs], n_classes, weights, size)
    names = [k for k, v in P.items() if v.requires_grad]
    gs = torch.autograd.grad(L["total"], [P[k] for k in names], allow_unused=True)
    return L, OrderedDict(zip(names, gs)), outputs


def sgd_step(P, grads, bufs, lr, momentum=0.9, weight_decay=1e-4):
    """torch.optim.SGD(momentum, weight_decay), dampening 0, no nesterov; params whose grad is
    None are skipped entirely (no decay, no momentum buffer)."""
    with torch.no_grad():
        for k, g in grads.items():
            if g is None:
                continue
            p = P[k]
            d = g + weight_decay * p
            if k not in bufs:
                bufs[k] = d.clone()
            else:
                bufs[k].mul_(momentum).add_(d)
            p.add_(bufs[k], alpha=-lr)


def poly_lr(base_lr, iter_num, max_iterations):
    """lr_ = base_lr * (1 - iter_num / max_iterations) ** 0.9 (train_..._BraTS.py:117)."""
    return base_lr * (1.0 - iter_num / max_iterations) ** 0.9


# --------------------------------------------------------------------------------------
# sliding-window inference + Dice metric (test_3D_BraTS.py:79-142, 175-187; val_3D.py:15-97)
# --------------------------------------------------------------------------------------


def window_starts(size, patch, stride):
    n = math.ceil((size - patch) / stride) + 1
    return [min(stride * i, size - patch) for i in range(n)]


def test_single_case(net_fn, image, stride_xy, stride_z, patch_size, num_classes):
    """Restatement of test_single_case: centred zero pad up to the patch size, window grid with
    the last window clamped to the border, softmax scores summed on the host in fp32 in x->y->z
    order, divided by the visit count, argmax over classes (first max on ties), un-pad.
    ``net_fn`` maps a [1,1,w,h,d] float tensor to logits [1,K,w,h,d]."""
    image = np.asarray(image)
    w, h, d = image.shape
    pads = []
    for s, p in zip((w, h, d), patch_size):
        tot = max(p - s, 0)
        pads.append((tot // 2, tot - tot // 2))
    padded = any(a + b > 0 for a, b in pads)
    if padded:
        image = np.pad(image, pads, mode="constant", constant_values=0)
    ww, hh, dd = image.shape
    score = np.zeros((num_classes, ww, hh, dd), dtype=np.float32)
    cnt = np.zeros((ww, hh, dd), dtype=np.float32)
    px, py, pz = patch_size
    for xs in window_starts(ww, px, stride_xy):
        for ys in window_starts(hh, py, stride_xy):
            for zs in window_starts(dd, pz, stride_z):
                patch = torch.from_numpy(np.ascontiguousarray(
                    image[xs:xs + px, ys:ys + py, zs:zs + pz][None, None].astype(np.float32)))
                with torch.no_grad():
                    prob = torch.softmax(net_fn(patch), dim=1)[0].cpu().numpy()
                score[:, xs:xs + px, ys:ys + py, zs:zs + pz] += prob
                cnt[xs:xs + px, ys:ys + py, zs:zs + pz] += 1
    score = score / cnt[None]
    label = np.argmax(score, axis=0)
    if padded:
        label = label[pads[0][0]:pads[0][0] + w, pads[1][0]:pads[1][0] + h, pads[2][0]:pads[2][0] + d]
    return label


def dice_metric(pred, gt):
    """calculate_metric_percase's Dice half (test_3D_BraTS.py:175-187) with MedPy 0.4.0
    ``metric.binary.dc`` restated: 2|A∩B| / (|A|+|B|) in float64; HD95 is out of scope.
    Returns (dice, counts) with counts = (|A∩B|, |A|, |B|) as exact integers."""
    a = np.asarray(pred) > 0
    b = np.asarray(gt) > 0
    inter = int(np.count_nonzero(a & b))
    na, nb = int(np.count_nonzero(a)), int(np.count_nonzero(b))
    if na > 0 and nb > 0:
        dice = 2.0 * inter / float(na + nb)
    elif na == 0 and nb == 0:
        dice = 1.0
    else:
        dice = 0.0
    return dice, (inter, na, nb)
