"""SwinUnet with the ICL heads — drop-in for the reference's networks/vision_transformer.py `SwinUnet` (constructor :33,
forward :90-108, load_from :111-147; SURVEY.md §8 row a20 / BASELINE config 4).

`config` is the reference's yacs-style attribute tree (config.DATA.IMG_SIZE, config.MODEL.SWIN.*, config.MODEL.DROP_RATE,
config.MODEL.DROP_PATH_RATE, config.TRAIN.USE_CHECKPOINT); any object with those attributes works, and
`swin_tiny_lite_config()` returns the effective values of configs/swin_tiny_patch4_window7_224_lite.yaml.
"""
import copy
from types import SimpleNamespace

import torch
import torch.nn as nn

from .. import functional as Fn
from .swinunet_icl import SwinTransformerSys
from .unet_icl import InherentConsistent as _InherentConsistent2d


def swin_tiny_lite_config(img_size=224, drop_path_rate=0.2, pretrain_ckpt=None):
    """configs/swin_tiny_patch4_window7_224_lite.yaml over the defaults of networks/config.py:29-102."""
    NS = SimpleNamespace
    return NS(DATA=NS(IMG_SIZE=img_size),
              MODEL=NS(DROP_RATE=0.0, DROP_PATH_RATE=drop_path_rate, PRETRAIN_CKPT=pretrain_ckpt,
                       SWIN=NS(PATCH_SIZE=4, IN_CHANS=3, EMBED_DIM=96, DEPTHS=[2, 2, 2, 2], DECODER_DEPTHS=[2, 2, 2, 1],
                               NUM_HEADS=[3, 6, 12, 24], WINDOW_SIZE=7, MLP_RATIO=4.0, QKV_BIAS=True, QK_SCALE=False, APE=False,
                               PATCH_NORM=True, FINAL_UPSAMPLE="expand_first")),
              TRAIN=NS(USE_CHECKPOINT=False))


class InherentConsistent(_InherentConsistent2d):
    """SSPA / USCL heads on decoder TOKEN tensors [B, N, C] (vision_transformer.py:185-264).  `proj_layers` / `norm_layers`
    exist as parameters but are bypassed (:246,258 are commented out in the reference), so they never receive a gradient."""

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
            tok = feats[i]
            q_in = next_Q if labeled else guided_Q[i].expand(BS, -1, -1)
            q, attn = self.class_decoders[i](q_in, tok, need_q)
            bs, K, H, N = attn.shape
            h = w = int(round(N ** 0.5))
            a = self.attn_convs0[i](attn.reshape(bs * K, H, 1, h, w))
            c1 = self.attn_convs1[i]
            feat_maps.append(Fn.planar_pointwise(a, c1.weight, c1.bias).reshape(bs, K, h, w))
            if need_q:
                qc = self.query_convs[i]
                next_Q = Fn.linear(q, qc.weight[:, :, 0], qc.bias)
                updated_Qs.append(Fn.batch_mean(q))
        Fn.flush_batch_counters()   # num_batches_tracked of the BatchNorms above: one multi-tensor add
        return feat_maps, updated_Qs


class SwinUnet(nn.Module):
    def __init__(self, config, img_size=224, num_classes=21843, zero_head=False, vis=False):
        super().__init__()
        self.num_classes, self.zero_head, self.config = num_classes, zero_head, config
        sw = config.MODEL.SWIN
        self.swin_unet = SwinTransformerSys(img_size=config.DATA.IMG_SIZE, patch_size=sw.PATCH_SIZE, in_chans=sw.IN_CHANS,
                                            num_classes=self.num_classes, embed_dim=sw.EMBED_DIM, depths=sw.DEPTHS, num_heads=sw.NUM_HEADS,
                                            window_size=sw.WINDOW_SIZE, mlp_ratio=sw.MLP_RATIO, qkv_bias=sw.QKV_BIAS, qk_scale=sw.QK_SCALE,
                                            drop_rate=config.MODEL.DROP_RATE, drop_path_rate=config.MODEL.DROP_PATH_RATE, ape=sw.APE,
                                            patch_norm=sw.PATCH_NORM, use_checkpoint=config.TRAIN.USE_CHECKPOINT)
        kw = dict(in_chans=(384, 192, 96), depths=(2, 2, 2), patch_size=sw.PATCH_SIZE, input_resolution=(14, 28, 56),
                  num_classes=self.num_classes, num_heads=(24, 12, 6), norm_layer=nn.LayerNorm)
        self.sspa = InherentConsistent(**kw)
        self.uscl = InherentConsistent(**kw)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def forward(self, x_lab, x_unlab=None, inference=False):
        if inference:
            if x_lab.size()[1] == 1:
                x_lab = x_lab.repeat(1, 3, 1, 1)
            return self.swin_unet(x_lab, inference=inference)
        if x_lab.size()[1] == 1 and x_unlab.size()[1] == 1:
            x_lab = x_lab.repeat(1, 3, 1, 1)
            x_unlab = x_unlab.repeat(1, 3, 1, 1)
            output_lab, output_unlab, feats_lab, feats_unlab = self.swin_unet(x_lab, x_unlab)
            feat_Maps_lab, updated_Qs_lab = self.sspa(feats_lab, None, "labeled")
            feat_Maps_consisunlab, _ = self.sspa(feats_unlab, None, "labeled")
            feat_Maps_unlab, _ = self.uscl(feats_unlab, updated_Qs_lab, "unlabeled", need_queries=False)
            return output_lab, output_unlab, feat_Maps_lab, feat_Maps_unlab, feat_Maps_consisunlab
        return None  # the reference falls through (returns None) for multi-channel training input (:96-108)

    def load_from(self, config):
        """Loads a Swin-T ImageNet checkpoint into the encoder and, mirrored, into the decoder (vision_transformer.py:111-147)."""
        path = config.MODEL.PRETRAIN_CKPT
        if path is None:
            print("none pretrain")
            return
        print("pretrained_path:{}".format(path))
        ckpt = torch.load(path, map_location="cuda" if torch.cuda.is_available() else "cpu")
        if "model" not in ckpt:
            ckpt = {k[17:]: v for k, v in ckpt.items() if "output" not in k}
            self.swin_unet.load_state_dict(ckpt, strict=False)
            return
        ckpt = ckpt["model"]
        own = self.swin_unet.state_dict()
        full = copy.deepcopy(ckpt)
        for k, v in ckpt.items():
            if "layers." in k:
                full["layers_up." + str(3 - int(k[7:8])) + k[8:]] = v
        for k in list(full.keys()):
            if k in own and full[k].shape != own[k].shape:
                del full[k]
        self.swin_unet.load_state_dict(full, strict=False)
