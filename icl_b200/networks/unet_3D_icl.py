"""unet_3D_icl — drop-in for the reference's networks/unet_3D_icl.py.

Module tree and parameter names are the reference's (so checkpoints interchange); every forward op is a
CUDA kernel from icl_b200 (backbone: backbone3d.Backbone3DFn; heads: icl_b200.functional)."""
import os
from typing import Sequence

import torch
import torch.nn as nn

from .. import functional as Fn
from .. import lanes as Ln
from .unet_3D import _Backbone3DModule


class DropPath(nn.Module):
    """Per-sample stochastic depth (MONAI 1.0.1 / timm semantics).  Only draws the per-sample scale
    r = Bernoulli(keep)/keep; the multiply is fused into Fn.add_scaled."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob
        self._queue = []  # test hook: replayed per-sample scales

    def sample(self, B, device):
        if self.drop_prob == 0.0 or not self.training:
            return None
        if self._queue:
            return self._queue.pop(0).to(device=device, dtype=torch.float32).reshape(B).contiguous()
        return _droppath_draw(1.0 - self.drop_prob, B, device)


# DropPath scales come from a pool drawn with ONE bernoulli_ + ONE div_ (instead of a pair of tiny kernels per draw: 36 draws per
# training step).  The pool holds as many rows as the previous refill period consumed (at least 16); it is refilled when exhausted
# or when CUDA-graph capture begins / ends (the draw must be a node of the graph that consumes it).
_DP_POOL = {}


def _droppath_draw(keep, B, device):
    capturing = torch.cuda.is_current_stream_capturing() if device.type == "cuda" else False
    key = (str(device), B, float(keep))
    st = _DP_POOL.get(key)
    if st is None or st["i"] >= st["pool"].shape[0] or st["cap"] != capturing:
        rows = 48 if st is None else max(16, st["pool"].shape[0])
        pool = torch.empty((rows, B), dtype=torch.float32, device=device).bernoulli_(keep).div_(keep)
        st = _DP_POOL[key] = {"pool": pool, "i": 0, "cap": capturing}
    r = st["pool"][st["i"]]
    st["i"] += 1
    return r


class MLP(nn.Module):
    """fc1 -> GELU(erf) -> fc2 (unet_3D_icl.py:299-315); GELU is fused into the fc1 GEMM epilogue."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        h = Fn.linear(x, self.fc1.weight, self.fc1.bias, act=1)
        return Fn.linear(h, self.fc2.weight, self.fc2.bias)


class Query_Attention(nn.Module):
    """Voxel -> class-proxy cross attention (unet_3D_icl.py:270-297)."""

    def __init__(self, dim, num_heads, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        if qk_scale is not None or attn_drop or proj_drop:
            raise NotImplementedError("icl_b200 Query_Attention implements the configuration the reference uses "
                                      "(qk_scale=None, attn_drop=proj_drop=0; unet_3D_icl.py:189-194)")
        self.dim, self.num_heads = dim, num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.fc_q = nn.Linear(dim, dim, bias=qkv_bias)
        self.fc_kv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, q, x, need_query=True):
        ql = Fn.linear(q, self.fc_q.weight, self.fc_q.bias)
        kv = Fn.linear(x, self.fc_kv.weight, self.fc_kv.bias)
        xv, attn = Fn.proxy_attention(ql, kv, self.num_heads, want_xv=need_query)
        out = Fn.linear(xv, self.proj.weight, self.proj.bias) if need_query else None
        return out, attn


class Class_Decoder(nn.Module):
    """unet_3D_icl.py:244-268.  norm3 / mlp2 act over the spatial axis N of the [B,K,H,N] attention map."""

    def __init__(self, dim, input_resolution, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.norm1_query = norm_layer(dim)
        self.attn = Query_Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = MLP(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        n = input_resolution[0] * input_resolution[1] * input_resolution[2]
        self.norm3 = norm_layer(n)
        self.mlp2 = MLP(in_features=n, hidden_features=n, act_layer=act_layer, drop=drop)

    def _r(self, B, device):
        return self.drop_path.sample(B, device) if isinstance(self.drop_path, DropPath) else None

    def draws(self, B, device):
        """The four DropPath scales of one call, drawn in the reference's order: q+dp(q), q+dp(mlp), a+dp(a), a+dp(mlp2)  (:264-267)."""
        return [self._r(B, device) for _ in range(4)]

    # The forward pass in independent pieces, so that InherentConsistent can put them on different lanes (icl_b200/lanes.py):
    # kv() and ql() feed the attention, q_branch() and attn_branch() consume its two outputs and do not see each other.
    def kv(self, feat):
        return Fn.linear(Fn.layer_norm(feat, self.norm1.weight, self.norm1.bias, self.norm1.eps), self.attn.fc_kv.weight, self.attn.fc_kv.bias)

    def ql(self, query):
        n = self.norm1_query
        return Fn.linear(Fn.layer_norm(query, n.weight, n.bias, n.eps), self.attn.fc_q.weight, self.attn.fc_q.bias)

    def q_branch(self, xv, rs):
        q = Fn.linear(xv, self.attn.proj.weight, self.attn.proj.bias)
        q = Fn.add_scaled(q, q, rs[0])
        return Fn.add_scaled(q, self.mlp(Fn.layer_norm(q, self.norm2.weight, self.norm2.bias, self.norm2.eps)), rs[1])

    def attn_branch(self, attn, rs):
        attn = Fn.add_scaled(attn, attn, rs[2])
        return Fn.add_scaled(attn, self.mlp2(Fn.layer_norm(attn, self.norm3.weight, self.norm3.bias, self.norm3.eps)), rs[3])

    def forward(self, query, feat, need_query=True):
        rs = self.draws(feat.shape[0], feat.device)
        xv, attn = Fn.proxy_attention(self.ql(query), self.kv(feat), self.attn.num_heads, want_xv=need_query)
        q = self.q_branch(xv, rs) if need_query else None
        return q, self.attn_branch(attn, rs)


class SeparableConv3d(nn.Module):
    """depthwise 3^3 -> BN3d -> ReLU -> pointwise -> BN3d -> ReLU (relu_first=False; unet_3D_icl.py:317-345)."""

    def __init__(self, inplanes, planes, kernel_size=(3, 3, 3), stride=(1, 1, 1), dilation=(1, 1, 1), relu_first=True, bias=False,
                 norm_layer=nn.BatchNorm3d):
        super().__init__()
        if relu_first or bias or tuple(kernel_size) != (3, 3, 3) or tuple(stride) != (1, 1, 1) or tuple(dilation) != (1, 1, 1):
            raise NotImplementedError("icl_b200 SeparableConv3d implements relu_first=False, bias=False, 3x3x3, stride/dilation 1")
        from collections import OrderedDict
        self.block = nn.Sequential(OrderedDict([
            ("depthwise", nn.Conv3d(inplanes, inplanes, kernel_size, stride=stride, padding=dilation, dilation=dilation, groups=inplanes,
                                    bias=False)),
            ("bn_depth", norm_layer(inplanes)),
            ("relu1", nn.ReLU(inplace=True)),
            ("pointwise", nn.Conv3d(inplanes, planes, (1, 1, 1), bias=False)),
            ("bn_point", norm_layer(planes)),
            ("relu2", nn.ReLU(inplace=True)),
        ]))

    def forward(self, x):
        b = self.block
        y = Fn.dwconv3d(x, b.depthwise.weight)
        y = Fn.bn_relu(y, b.bn_depth, b.bn_depth.training)
        y = Fn.planar_pointwise(y, b.pointwise.weight, None)
        return Fn.bn_relu(y, b.bn_point, b.bn_point.training)


class InherentConsistent(nn.Module):
    """SSPA / USCL heads (unet_3D_icl.py:155-242)."""

    def __init__(self, in_chans: Sequence[int], depths: Sequence[int], patch_size: Sequence[int], input_resolution: Sequence[int],
                 num_classes: int, num_heads: Sequence[int], norm_layer=nn.LayerNorm, patch_norm: bool = False, spatial_dims: int = 3,
                 drop_path_rate: float = 0.1):
        super().__init__()
        self.in_chans, self.patch_size, self.patch_norm, self.depth = in_chans, patch_size, patch_norm, depths
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        self.proj_layers = nn.ModuleList()
        self.norm_layers = nn.ModuleList()
        self.class_decoders = nn.ModuleList()
        self.attn_convs0 = nn.ModuleList()
        self.attn_convs1 = nn.ModuleList()
        self.query_convs = nn.ModuleList()
        for i in range(len(depths)):
            r = input_resolution[i]
            self.proj_layers.append(nn.Conv3d(in_chans[i], in_chans[i], kernel_size=(1, 1, 1), stride=(1, 1, 1)))
            self.norm_layers.append(norm_layer(in_chans[i]))
            self.class_decoders.append(Class_Decoder(dim=in_chans[i], input_resolution=(r, r, r), num_heads=num_heads[i], mlp_ratio=4.0,
                                                     qkv_bias=True, qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=dpr[1],
                                                     norm_layer=norm_layer))
            self.attn_convs0.append(SeparableConv3d(num_heads[i], num_heads[i], (3, 3, 3), norm_layer=nn.BatchNorm3d, relu_first=False))
            self.attn_convs1.append(nn.Conv3d(num_heads[i], 1, kernel_size=(1, 1, 1), stride=(1, 1, 1)))
            self.query_convs.append(nn.Conv1d(in_chans[i], in_chans[i] // 2, kernel_size=1, stride=1, padding=0))
        self.guided_Q = nn.Parameter(torch.zeros(1, num_classes, in_chans[0]))

    def forward(self, feats, guided_Q=None, modal="labeled", need_queries=True, lanes=None):
        """`need_queries=False` (our extension, default keeps reference behaviour) skips the softmax.V / proj /
        mlp / query_convs branch when the caller discards updated_Qs in 'unlabeled' mode — that branch is
        dead in the reference's graph (unet_3D_icl.py:147) and its parameters keep grad None.

        `lanes` (our extension): {"main", "q", "s": [one stream per scale]} from icl_b200.lanes — the token / attention-map work of
        scale i is enqueued on lanes["s"][i], the query chain on lanes["q"]; the returned tensors are then NOT yet ordered with
        respect to the caller's stream: the caller joins the lanes (unet_3D_icl.forward does).  None: everything on the current stream."""
        feat_maps, updated_Qs = [], []
        BS = feats[0].shape[0]
        if modal not in ("labeled", "unlabeled"):
            return feat_maps, updated_Qs
        Fn.defer_batch_counters()
        labeled = modal == "labeled"
        need_q = need_queries or labeled
        main = lanes["main"] if lanes else None
        lq = lanes["q"] if lanes else None
        # every DropPath scale of this pass first, in the reference's order (scale by scale), on the caller's stream
        draws = [self.class_decoders[i].draws(BS, feats[i].device) for i in range(len(self.depth))]
        next_Q = self.guided_Q.expand(BS, -1, -1) if labeled else None
        Ln.handoff(lq, main, *[r for rs in draws for r in rs[:2]])
        for i in range(len(self.depth)):
            f = feats[i]
            B, C, d, h, w = f.shape
            ls = lanes["s"][i] if lanes else None
            dec, rs = self.class_decoders[i], draws[i]
            Ln.handoff(ls, main, f, *rs[2:])
            with Ln.on(ls):
                pl = self.proj_layers[i]
                # 1x1x1 conv on channels-last rows == Linear; flatten(2).transpose(1,2) is free in NDHWC
                tok = Fn.linear(f.permute(0, 2, 3, 4, 1).reshape(B, d * h * w, C), pl.weight.reshape(C, C), pl.bias)
                nl = self.norm_layers[i]
                kv = dec.kv(Fn.layer_norm(tok, nl.weight, nl.bias, nl.eps))
            with Ln.on(lq):
                ql = dec.ql(next_Q if labeled else guided_Q[i].expand(BS, -1, -1))
            Ln.handoff(ls, lq, ql)
            with Ln.on(ls):
                xv, attn = Fn.proxy_attention(ql, kv, dec.attn.num_heads, want_xv=need_q)
            if need_q:
                Ln.handoff(lq, ls, xv)
                with Ln.on(lq):
                    q = dec.q_branch(xv, rs)
                    qc = self.query_convs[i]
                    next_Q = Fn.linear(q, qc.weight[:, :, 0], qc.bias)
                    updated_Qs.append(Fn.batch_mean(q))
            with Ln.on(ls):
                attn = dec.attn_branch(attn, rs)
                bs, K, H, N = attn.shape
                a = attn.reshape(bs * K, H, d, h, w)
                a = self.attn_convs0[i](a)
                c1 = self.attn_convs1[i]
                feat_maps.append(Fn.planar_pointwise(a, c1.weight, c1.bias).reshape(bs, K, d, h, w))
        Fn.flush_batch_counters()   # num_batches_tracked of the BatchNorms above: one multi-tensor add
        return feat_maps, updated_Qs


class unet_3D_icl(_Backbone3DModule):
    def __init__(self, feature_scale=4, n_classes=21, is_deconv=True, in_channels=3, is_batchnorm=True):
        super().__init__(feature_scale, n_classes, is_deconv, in_channels, is_batchnorm)
        f = self.filters
        icl_in_chans = (f[4], f[3], f[2])
        icl_in_resolutions = [6, 12, 24]
        kw = dict(in_chans=icl_in_chans, depths=(2, 2, 2), patch_size=(2, 2, 2), input_resolution=icl_in_resolutions,
                  num_classes=n_classes, num_heads=(16, 8, 4), norm_layer=nn.LayerNorm)
        self.sspa = InherentConsistent(**kw)
        self.uscl = InherentConsistent(**kw)

    def forward(self, x_lab, x_unlab=None, inference=None):
        if inference:
            return self._run(x_lab)[0]
        dev = x_lab.device
        pair = os.environ.get("ICL_DISABLE_PAIR") != "1" and x_lab.shape[1:] == x_unlab.shape[1:]
        if not (pair and Ln.enabled(dev)):
            if pair:
                # both passes share every weight and InstanceNorm is per sample: one batched pass, same values
                (final_lab, center_lab, up4_lab, up3_lab), (final_unlab, center_unlab, up4_unlab, up3_unlab) = self._run_pair(x_lab, x_unlab)
            else:
                final_lab, center_lab, up4_lab, up3_lab = self._run(x_lab)
                final_unlab, center_unlab, up4_unlab, up3_unlab = self._run(x_unlab)
            feats_lab = [center_lab, up4_lab, up3_lab]
            feats_unlab = [center_unlab, up4_unlab, up3_unlab]
            feat_Maps_lab, updated_Qs_lab = self.sspa(feats_lab, "labeled")
            feat_Maps_consis, _ = self.sspa(feats_unlab, "labeled")
            feat_Maps_unlab, _ = self.uscl(feats_unlab, updated_Qs_lab, "unlabeled", need_queries=False)
            return final_lab, final_unlab, feat_Maps_lab, feat_Maps_unlab, feat_Maps_consis
        # Lanes (icl_b200/lanes.py).  The batched backbone pass is two autograd nodes: the heads only read center / up4 / up3, so they
        # are enqueued on side streams right after the lower node and run next to up_concat2 / up_concat1 / final (forward) and next
        # to those layers' backward.  Per-scale lanes s0..s2 are shared by the two sspa passes (same BatchNorm modules: same lane, so
        # their running statistics are updated in the reference's order), the query chain has a lane, and the uscl pass, which only
        # waits for the labeled pass's updated proxies, has lanes of its own.
        feats_lab, feats_unlab, top = self._run_pair_low(x_lab, x_unlab)
        L = Ln.get(dev, ["q", "s0", "s1", "s2", "uq", "u0", "u1", "u2"])
        main = L["main"]
        la = {"main": main, "q": L["q"], "s": [L["s0"], L["s1"], L["s2"]]}
        lu = {"main": main, "q": L["uq"], "s": [L["u0"], L["u1"], L["u2"]]}
        feat_Maps_lab, updated_Qs_lab = self.sspa(list(feats_lab), "labeled", lanes=la)
        Ln.handoff(L["uq"], L["q"], *updated_Qs_lab)
        feat_Maps_consis, _ = self.sspa(list(feats_unlab), "labeled", lanes=la)
        feat_Maps_unlab, _ = self.uscl(list(feats_unlab), updated_Qs_lab, "unlabeled", need_queries=False, lanes=lu)
        final_lab, final_unlab = self._run_pair_top(top)
        for k, s in L.items():
            if k != "main":
                Ln.handoff(main, s)
        for t in feat_Maps_lab + feat_Maps_unlab + feat_Maps_consis:
            t.record_stream(main)
        return final_lab, final_unlab, feat_Maps_lab, feat_Maps_unlab, feat_Maps_consis

    @staticmethod
    def apply_argmax_softmax(pred):
        return torch.softmax(pred, dim=1)
