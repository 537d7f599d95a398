"""UNet_icl — drop-in for the reference's networks/unet_icl.py (constructor :197, forward :237-252): the 2D U-Net backbone
run on the labeled and the unlabeled batch plus the SSPA / USCL Inherent-Consistent-Learning heads at 32^2 / 64^2 / 128^2.

Module tree and parameter names are the reference's.  The token-side classes (Class_Decoder, Query_Attention, MLP) are the
dimension-agnostic ones of unet_3D_icl.py; the 2D convolutions of the heads run through the planar depth-1 forms of the
3D kernels (icl_b200.functional)."""
from collections import OrderedDict
from typing import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as Fn
from .unet import Decoder, Encoder
from .unet_3D_icl import Class_Decoder


class SeparableConv2d(nn.Module):
    """depthwise 3x3 -> BN2d -> ReLU -> pointwise 1x1 -> BN2d -> ReLU (relu_first=False; unet_icl.py:98-126)."""

    def __init__(self, inplanes, planes, kernel_size=3, stride=1, dilation=1, relu_first=True, bias=False, norm_layer=nn.BatchNorm2d):
        super().__init__()
        ks = tuple(kernel_size) if isinstance(kernel_size, (tuple, list)) else (kernel_size, kernel_size)
        if relu_first or bias or ks != (3, 3) or stride != 1 or dilation != 1:
            raise NotImplementedError("icl_b200 SeparableConv2d implements relu_first=False, bias=False, 3x3, stride/dilation 1")
        self.block = nn.Sequential(OrderedDict([
            ("depthwise", nn.Conv2d(inplanes, inplanes, 3, stride=1, padding=1, dilation=1, groups=inplanes, bias=False)),
            ("bn_depth", norm_layer(inplanes)),
            ("relu1", nn.ReLU(inplace=True)),
            ("pointwise", nn.Conv2d(inplanes, planes, 1, bias=False)),
            ("bn_point", norm_layer(planes)),
            ("relu2", nn.ReLU(inplace=True)),
        ]))

    def forward(self, x):
        """x: planar [NB, CH, 1, h, w]."""
        b = self.block
        w3 = F.pad(b.depthwise.weight.unsqueeze(2), (0, 0, 0, 0, 1, 1))  # 2D taps in the centre depth plane
        y = Fn.dwconv3d(x, w3)
        y = Fn.bn_relu(y, b.bn_depth, b.bn_depth.training)
        y = Fn.planar_pointwise(y, b.pointwise.weight, None)
        return Fn.bn_relu(y, b.bn_point, b.bn_point.training)


class InherentConsistent(nn.Module):
    """SSPA / USCL heads, spatial_dims = 2 (unet_icl.py:254-343)."""

    def __init__(self, in_chans: Sequence[int], depths: Sequence[int], patch_size: Sequence[int], input_resolution: Sequence[int],
                 num_classes: int, num_heads: Sequence[int], norm_layer=nn.LayerNorm, patch_norm: bool = False, spatial_dims: int = 2,
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
            self.proj_layers.append(nn.Conv2d(in_chans[i], in_chans[i], kernel_size=(1, 1), stride=(1, 1)))
            self.norm_layers.append(norm_layer(in_chans[i]))
            self.class_decoders.append(Class_Decoder(dim=in_chans[i], input_resolution=(r, r, 1), num_heads=num_heads[i], mlp_ratio=4.0,
                                                     qkv_bias=True, qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=dpr[1],
                                                     norm_layer=norm_layer))
            self.attn_convs0.append(SeparableConv2d(num_heads[i], num_heads[i], (3, 3), norm_layer=nn.BatchNorm2d, relu_first=False))
            self.attn_convs1.append(nn.Conv2d(num_heads[i], 1, kernel_size=(1, 1), stride=(1, 1)))
            self.query_convs.append(nn.Conv1d(in_chans[i], in_chans[i] // 2, kernel_size=1, stride=1, padding=0))
        self.guided_Q = nn.Parameter(torch.zeros(1, num_classes, in_chans[0]))

    def forward(self, feats, guided_Q=None, modal="labeled", need_queries=True):
        feat_maps, updated_Qs = [], []
        BS = feats[0].shape[0]
        if modal not in ("labeled", "unlabeled"):
            return feat_maps, updated_Qs
        Fn.defer_batch_counters()
        labeled = modal == "labeled"
        need_q = need_queries or labeled
        next_Q = self.guided_Q.expand(BS, -1, -1) if labeled else None
        for i in range(len(self.depth)):
            f = feats[i]
            B, C, h, w = f.shape
            pl = self.proj_layers[i]
            tok = Fn.linear(f.permute(0, 2, 3, 1).reshape(B, h * w, C), pl.weight.reshape(C, C), pl.bias)
            nl = self.norm_layers[i]
            tok = Fn.layer_norm(tok, nl.weight, nl.bias, nl.eps)
            q_in = next_Q if labeled else guided_Q[i].expand(BS, -1, -1)
            q, attn = self.class_decoders[i](q_in, tok, need_q)
            bs, K, H, N = attn.shape
            a = self.attn_convs0[i](attn.reshape(bs * K, H, 1, h, w))
            c1 = self.attn_convs1[i]
            feat_maps.append(Fn.planar_pointwise(a, c1.weight, c1.bias).reshape(bs, K, h, w))
            if need_q:
                qc = self.query_convs[i]
                next_Q = Fn.linear(q, qc.weight[:, :, 0], qc.bias)
                updated_Qs.append(Fn.batch_mean(q))
        Fn.flush_batch_counters()   # num_batches_tracked of the BatchNorms above: one multi-tensor add
        return feat_maps, updated_Qs


class UNet_icl(nn.Module):
    def __init__(self, in_chns, class_num):
        super().__init__()
        params = {"in_chns": in_chns, "feature_chns": [16, 32, 64, 128, 256], "input_resolution": [16, 32, 64, 128, 256],
                  "num_heads": (2, 4, 8), "depths": (2, 2, 2), "dropout": [0.05, 0.1, 0.2, 0.3, 0.5], "class_num": class_num,
                  "bilinear": False, "acti_func": "relu"}
        self.encoder = Encoder(params)
        self.decoder = Decoder(params, return_feats=True)
        f, r = params["feature_chns"], params["input_resolution"]
        kw = dict(in_chans=(f[3], f[2], f[1]), depths=params["depths"], patch_size=(2, 2), input_resolution=(r[1], r[2], r[3]),
                  num_classes=class_num, num_heads=params["num_heads"][::-1], norm_layer=nn.LayerNorm)
        self.sspa = InherentConsistent(**kw)
        self.uscl = InherentConsistent(**kw)

    def forward(self, x_lab, x_unlab=None, inference=False):
        output_lab, feats_lab = self.decoder(self.encoder(x_lab))
        if inference:
            return output_lab
        output_unlab, feats_unlab = self.decoder(self.encoder(x_unlab))
        feat_Maps_lab, updated_Qs_lab = self.sspa(feats_lab, "labeled")
        feat_Maps_consisunlab, _ = self.sspa(feats_unlab, "labeled")
        feat_Maps_unlab, _ = self.uscl(feats_unlab, updated_Qs_lab, "unlabeled", need_queries=False)
        return output_lab, output_unlab, feat_Maps_lab, feat_Maps_unlab, feat_Maps_consisunlab
