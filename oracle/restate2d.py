"""CPU restatement of the 2D ICL path (TEST INFRASTRUCTURE — see oracle/__init__.py): UNet / UNet_icl of
networks/unet_icl.py (identical blocks in networks/unet.py) and the 2D loss callables of utils/losses.py.

Same functional form as oracle/restate.py: a flat ``state_dict``-style mapping ``P`` (name -> tensor), every function
citing the reference lines it follows (paths relative to /root/reference/code).  The token-side ICL classes
(Class_Decoder, Query_Attention, MLP) are dimension-agnostic and shared with restate.py.
Pinned against the live reference by tests/test_oracle_vs_reference.py and by the fixtures of oracle/make_golden.py.
"""
import torch
import torch.nn.functional as F

from . import restate as R

DROPOUT_2D = (0.05, 0.1, 0.2, 0.3, 0.5)  # unet_icl.py:205
ICL_HEADS_2D = (8, 4, 2)                 # num_heads (2, 4, 8) reversed, unet_icl.py:203,215


def _bn(x, P, prefix, training, momentum=0.1, eps=1e-5):
    """nn.BatchNorm2d: batch statistics + running-stat update in training, running statistics in eval."""
    return F.batch_norm(x, P.get(prefix + ".running_mean"), P.get(prefix + ".running_var"), P[prefix + ".weight"], P[prefix + ".bias"],
                        training, momentum, eps)


def conv_block(P, prefix, x, p_drop, rand, training):
    """ConvBlock.forward (unet_icl.py:41-57): Conv2d 3x3 p1 -> BatchNorm2d -> LeakyReLU(0.01) -> Dropout(p) -> Conv2d -> BN -> LeakyReLU."""
    q = prefix + ".conv_conv"
    y = F.conv2d(x, P[q + ".0.weight"], P[q + ".0.bias"], padding=1)
    y = F.leaky_relu(_bn(y, P, q + ".1", training), 0.01)
    y = rand.dropout(y, p_drop) if p_drop > 0 else y
    y = F.conv2d(y, P[q + ".4.weight"], P[q + ".4.bias"], padding=1)
    return F.leaky_relu(_bn(y, P, q + ".5", training), 0.01)


def encoder(P, x, rand, training):
    """Encoder.forward (unet_icl.py:149-155): in_conv, then 4 x (MaxPool2d(2) -> ConvBlock)."""
    x0 = conv_block(P, "encoder.in_conv", x, DROPOUT_2D[0], rand, training)
    feats = [x0]
    for i in range(1, 5):
        feats.append(conv_block(P, "encoder.down%d.maxpool_conv.1" % i, F.max_pool2d(feats[-1], 2), DROPOUT_2D[i], rand, training))
    return feats


def up_block(P, prefix, x1, x2, rand, training):
    """UpBlock.forward with bilinear=True (the default the Decoder uses, unet_icl.py:79,90-96): conv1x1 -> bilinear x2
    (align_corners=True) -> cat([skip, up], 1) -> ConvBlock(dropout 0)."""
    x1 = F.conv2d(x1, P[prefix + ".conv1x1.weight"], P[prefix + ".conv1x1.bias"])
    x1 = F.interpolate(x1, scale_factor=2, mode="bilinear", align_corners=True)
    return conv_block(P, prefix + ".conv", torch.cat([x2, x1], 1), 0.0, rand, training)


def decoder(P, feats, rand, training):
    """Decoder.forward (unet_icl.py:180-194): returns (out_conv logits, [x_1, x_2, x_3])."""
    x0, x1, x2, x3, x4 = feats
    u1 = up_block(P, "decoder.up1", x4, x3, rand, training)
    u2 = up_block(P, "decoder.up2", u1, x2, rand, training)
    u3 = up_block(P, "decoder.up3", u2, x1, rand, training)
    u4 = up_block(P, "decoder.up4", u3, x0, rand, training)
    out = F.conv2d(u4, P["decoder.out_conv.weight"], P["decoder.out_conv.bias"], padding=1)
    return out, [u1, u2, u3]


def unet2d_forward(P, x, rand=None, training=True):
    """UNet.forward (unet.py:318-321) / the inference branch of UNet_icl.forward (unet_icl.py:238-242)."""
    rand = rand or R.NoRand()
    return decoder(P, encoder(P, x, rand, training), rand, training)[0]


def separable_conv2d(P, prefix, x, training):
    """SeparableConv2d(relu_first=False).forward (unet_icl.py:98-126)."""
    C = x.shape[1]
    y = F.conv2d(x, P[prefix + ".block.depthwise.weight"], None, 1, 1, 1, groups=C)
    y = F.relu(_bn(y, P, prefix + ".block.bn_depth", training))
    y = F.conv2d(y, P[prefix + ".block.pointwise.weight"], None)
    return F.relu(_bn(y, P, prefix + ".block.bn_point", training))


def inherent_consistent_2d(P, prefix, feats, guided_Q=None, modal="labeled", heads=ICL_HEADS_2D, rand=None, training=True):
    """InherentConsistent.forward, spatial_dims=2 (unet_icl.py:300-343)."""
    rand = rand or R.NoRand()
    feat_maps, updated_Qs = [], []
    B = feats[0].shape[0]
    next_Q = P[prefix + ".guided_Q"].expand(B, -1, -1) if modal == "labeled" else None
    for i, f in enumerate(feats):
        h, w = f.shape[2:]
        tok = F.conv2d(f, P["%s.proj_layers.%d.weight" % (prefix, i)], P["%s.proj_layers.%d.bias" % (prefix, i)])
        tok = R._ln(tok.flatten(2).transpose(1, 2), P, "%s.norm_layers.%d" % (prefix, i))
        q_in = next_Q if modal == "labeled" else guided_Q[i].expand(B, -1, -1)
        q, a = R.class_decoder(P, "%s.class_decoders.%d" % (prefix, i), q_in, tok, heads[i], rand)
        bs, K, H, N = a.shape
        a = a.contiguous().view(bs * K, H, h, w)
        a = separable_conv2d(P, "%s.attn_convs0.%d" % (prefix, i), a, training)
        fm = F.conv2d(a, P["%s.attn_convs1.%d.weight" % (prefix, i)], P["%s.attn_convs1.%d.bias" % (prefix, i)])
        feat_maps.append(fm.reshape(bs, K, h, w))
        wq = P["%s.query_convs.%d.weight" % (prefix, i)]
        next_Q = F.linear(q, wq[:, :, 0], P["%s.query_convs.%d.bias" % (prefix, i)])
        updated_Qs.append(q.mean(dim=0, keepdim=True))
    return feat_maps, updated_Qs


def unet_icl_forward(P, x_lab, x_unlab=None, inference=False, rand=None, training=True):
    """UNet_icl.forward (unet_icl.py:237-252): two encoder/decoder passes (BatchNorm statistics per branch), then
    sspa(lab), sspa(unlab), uscl(unlab, queries of the labeled pass)."""
    rand = rand or R.NoRand()
    out_lab, feats_lab = decoder(P, encoder(P, x_lab, rand, training), rand, training)
    if inference:
        return out_lab
    out_unlab, feats_unlab = decoder(P, encoder(P, x_unlab, rand, training), rand, training)
    maps_lab, Qs_lab = inherent_consistent_2d(P, "sspa", feats_lab, None, "labeled", rand=rand, training=training)
    maps_consis, _ = inherent_consistent_2d(P, "sspa", feats_unlab, None, "labeled", rand=rand, training=training)
    maps_unlab, _ = inherent_consistent_2d(P, "uscl", feats_unlab, Qs_lab, "unlabeled", rand=rand, training=training)
    return out_lab, out_unlab, maps_lab, maps_unlab, maps_consis


# ---------------------------------------------------------------------------------------------------------- losses
def aux_loss_2d(feat_maps, labels, n_classes, size):
    """AuxLoss.forward (utils/losses.py:241-251): bilinear (align_corners=False) resize, CE + DiceLoss(softmax=True), mean over maps."""
    ce, dice = 0.0, 0.0
    for fm in feat_maps:
        up = F.interpolate(fm.float(), size=list(size), mode="bilinear")
        ce = ce + R.ce_loss(up, labels.long())
        dice = dice + R.dice_loss(up, labels.unsqueeze(1), n_classes, softmax=True)
    return ce / len(feat_maps) + dice / len(feat_maps)


def pseudo_soft_loss_2d(feat_maps, predicts, size):
    """PseudoSoftLoss.forward (utils/losses.py:277-285)."""
    tgt = predicts.clone().detach()
    tot = 0.0
    for fm in feat_maps:
        tot = tot + R.softmax_dice(F.interpolate(fm.float(), size=list(size), mode="bilinear"), tgt)
    return tot / len(feat_maps)


WEIGHTS_2D = (1.0, 1.0, 1.0, 1.0, 50.0)  # ce, dice, aux, pse, cons: train_inherent_consistent_unet_2D.py:126-127


def icl_losses_2d(outputs, labels_lab, n_classes, weights=WEIGHTS_2D):
    """The five loss terms of the 2D loop (train_inherent_consistent_unet_2D.py:119-127)."""
    size = tuple(outputs[0].shape[2:])
    L = {}
    L["ce"] = R.ce_loss(outputs[0], labels_lab.long())
    L["dice"] = R.dice_loss(outputs[0], labels_lab.unsqueeze(1), n_classes, softmax=True)
    L["aux"] = aux_loss_2d(outputs[2], labels_lab, n_classes, size)
    L["pse"] = pseudo_soft_loss_2d(outputs[3], outputs[1], size)
    L["cons"] = R.softmax_mse_loss(outputs[3], outputs[4])
    w = weights
    L["total"] = w[0] * L["ce"] + w[1] * L["dice"] + w[2] * L["aux"] + w[3] * L["pse"] + w[4] * L["cons"]
    return L
