"""CPU restatement of the Swin-UNet ICL path (TEST INFRASTRUCTURE — see oracle/__init__.py): SwinTransformerSys of
networks/swinunet_icl.py and SwinUnet + the token-input ICL heads of networks/vision_transformer.py (BASELINE config 4).

Same functional form as oracle/restate.py: a flat ``state_dict``-style mapping ``P`` (name -> tensor), every function citing the
reference lines it follows (paths relative to /root/reference/code).  The formulation is deliberately NOT the reference's
(no roll / window_partition / window_reverse / mask buffers): windows are gathered with one index table built from
coordinates, the shift mask comes from wrap-around region ids, and patch merging / expanding are reshapes — so agreeing with
the live reference (tests/test_oracle_vs_reference.py) is a real check of both.
Pinned against the live reference and by the fixtures of oracle/make_golden.py (swin_sys_mini, step_cfg4).
"""
import torch
import torch.nn.functional as F

from . import restate as R
from . import restate2d as R2

SWIN_TINY_LITE = dict(img_size=224, patch_size=4, in_chans=3, embed_dim=96, depths=(2, 2, 2, 2), num_heads=(3, 6, 12, 24), window_size=7,
                      drop_path_rate=0.2)   # configs/swin_tiny_patch4_window7_224_lite.yaml over networks/config.py:29-102
ICL_HEADS_SWIN = (24, 12, 6)                # vision_transformer.py:57


def window_token_index(H, W, ws, shift):
    """[nW, ws*ws] original token position of every slot of every window of the map rolled by -shift
    (roll :262-266, window_partition :34-48 of swinunet_icl.py)."""
    oy = (torch.arange(H) + shift) % H
    ox = (torch.arange(W) + shift) % W
    pos = oy[:, None] * W + ox[None, :]
    return pos.view(H // ws, ws, W // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)


def shift_mask(H, W, ws, shift):
    """[nW, N, N] additive mask: -100 between tokens from different wrap-around regions (swinunet_icl.py:217-245)."""
    def rid(n):
        c = torch.arange(n)
        return (c >= n - ws).long() + (c >= n - shift).long()
    r = (rid(H)[:, None] * 3 + rid(W)[None, :]).view(H // ws, ws, W // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    return torch.where(r[:, :, None] != r[:, None, :], -100.0, 0.0)


def relative_position_index(ws):
    """swinunet_icl.py:88-104."""
    t = torch.arange(ws * ws)
    y, x = t // ws, t % ws
    return (y[:, None] - y[None, :] + ws - 1) * (2 * ws - 1) + (x[:, None] - x[None, :] + ws - 1)


def window_attention_tokens(qkv, table, H, W, num_heads, ws, shift):
    """WindowAttention.forward without the qkv / proj Linears (swinunet_icl.py:127-152) applied to token-major qkv [B, H*W, 3C]
    with the shift / partition / reverse of SwinTransformerBlock.forward (:258-287) expressed as one gather and one scatter."""
    B, L, C3 = qkv.shape
    C = C3 // 3
    hd = C // num_heads
    pos = window_token_index(H, W, ws, shift).to(qkv.device)
    nW, N = pos.shape
    x = qkv[:, pos.reshape(-1)].view(B, nW, N, 3, num_heads, hd).permute(3, 0, 1, 4, 2, 5)   # [3, B, nW, nH, N, hd]
    q, k, v = x[0] * hd ** -0.5, x[1], x[2]
    attn = q @ k.transpose(-2, -1)
    bias = table[relative_position_index(ws).reshape(-1).to(qkv.device)].view(N, N, num_heads).permute(2, 0, 1)
    attn = attn + bias[None, None]
    if shift > 0:
        attn = attn + shift_mask(H, W, ws, shift).to(qkv.device)[None, :, None]
    out = (attn.softmax(-1) @ v).permute(0, 1, 3, 2, 4).reshape(B, nW * N, C)
    inv = torch.argsort(pos.reshape(-1))
    return out[:, inv]


def swin_block(P, prefix, x, res, num_heads, ws, shift, rand, p_drop=0.0):
    """SwinTransformerBlock.forward (swinunet_icl.py:249-293); DropPath draws: attention branch, then MLP branch."""
    H, W = res
    if min(res) <= ws:
        ws, shift = min(res), 0
    y = R._ln(x, P, prefix + ".norm1")
    qkv = F.linear(y, P[prefix + ".attn.qkv.weight"], P[prefix + ".attn.qkv.bias"])
    y = window_attention_tokens(qkv, P[prefix + ".attn.relative_position_bias_table"], H, W, num_heads, ws, shift)
    y = F.linear(y, P[prefix + ".attn.proj.weight"], P[prefix + ".attn.proj.bias"])
    x = x + rand.droppath(y, p_drop)
    return x + rand.droppath(R._mlp(R._ln(x, P, prefix + ".norm2"), P, prefix + ".mlp"), p_drop)


def patch_merging(P, prefix, x, res):
    """PatchMerging.forward (swinunet_icl.py:330-351)."""
    H, W = res
    B, L, C = x.shape
    x = x.view(B, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 4, 2, 5).reshape(B, L // 4, 4 * C)
    return F.linear(R._ln(x, P, prefix + ".norm"), P[prefix + ".reduction.weight"])


def _shuffle(x, H, W, p):
    B, L, C = x.shape
    return x.view(B, H, W, p, p, C // (p * p)).permute(0, 1, 3, 2, 4, 5).reshape(B, L * p * p, C // (p * p))


def patch_expand(P, prefix, x, res, p=2):
    """PatchExpand.forward (:372-387) / FinalPatchExpand_X4.forward (:400-415): Linear -> pixel shuffle of tokens -> LayerNorm."""
    x = F.linear(x, P[prefix + ".expand.weight"])
    return R._ln(_shuffle(x, res[0], res[1], p), P, prefix + ".norm")


def swin_sys_branch(P, x, cfg, rand, prefix=""):
    """forward_features + forward_up_features + up_x4 of SwinTransformerSys (swinunet_icl.py:752-794) for one image batch.
    Returns (logits [B,K,H,W], [decoder token maps at 1/16, 1/8, 1/4 resolution])."""
    ps, E, depths, heads, ws = cfg["patch_size"], cfg["embed_dim"], cfg["depths"], cfg["num_heads"], cfg["window_size"]
    nl = len(depths)
    dpr = torch.linspace(0, cfg["drop_path_rate"], sum(depths)).tolist()   # stochastic depth decay rule (:650-651)
    B, Cin, Hi, Wi = x.shape
    r0 = (Hi // ps, Wi // ps)
    # PatchEmbed (:586-594)
    t = F.conv2d(x, P[prefix + "patch_embed.proj.weight"], P[prefix + "patch_embed.proj.bias"], stride=ps).flatten(2).transpose(1, 2)
    t = R._ln(t, P, prefix + "patch_embed.norm")
    skips = []
    for i in range(nl):
        res = (r0[0] >> i, r0[1] >> i)
        skips.append(t)
        for b in range(depths[i]):
            t = swin_block(P, "%slayers.%d.blocks.%d" % (prefix, i, b), t, res, heads[i], ws, 0 if b % 2 == 0 else ws // 2, rand,
                           dpr[sum(depths[:i]) + b])
        if i < nl - 1:
            t = patch_merging(P, "%slayers.%d.downsample" % (prefix, i), t, res)
    t = R._ln(t, P, prefix + "norm")
    feats = []
    for i in range(nl):
        j = nl - 1 - i
        res = (r0[0] >> j, r0[1] >> j)
        if i == 0:
            t = patch_expand(P, prefix + "layers_up.0", t, res)
            continue
        t = F.linear(torch.cat([t, skips[j]], -1), P["%sconcat_back_dim.%d.weight" % (prefix, i)], P["%sconcat_back_dim.%d.bias" % (prefix, i)])
        for b in range(depths[j]):
            t = swin_block(P, "%slayers_up.%d.blocks.%d" % (prefix, i, b), t, res, heads[j], ws, 0 if b % 2 == 0 else ws // 2, rand,
                           dpr[sum(depths[:j]) + b])
        feats.append(t)
        if i < nl - 1:
            t = patch_expand(P, "%slayers_up.%d.upsample" % (prefix, i), t, res)
    t = R._ln(t, P, prefix + "norm_up")
    t = patch_expand(P, prefix + "up", t, r0, 4)
    logits = F.linear(t, P[prefix + "output.weight"].flatten(1)).view(B, 4 * r0[0], 4 * r0[1], -1).permute(0, 3, 1, 2)
    return logits, feats


def inherent_consistent_tokens(P, prefix, feats, guided_Q=None, modal="labeled", heads=ICL_HEADS_SWIN, rand=None, training=True):
    """InherentConsistent.forward of vision_transformer.py:232-264: token tensors go straight into the class decoders
    (proj_layers / norm_layers bypassed, :246,258)."""
    rand = rand or R.NoRand()
    feat_maps, updated_Qs = [], []
    B = feats[0].shape[0]
    next_Q = P[prefix + ".guided_Q"].expand(B, -1, -1) if modal == "labeled" else None
    for i, tok in enumerate(feats):
        q_in = next_Q if modal == "labeled" else guided_Q[i].expand(B, -1, -1)
        q, a = R.class_decoder(P, "%s.class_decoders.%d" % (prefix, i), q_in, tok, heads[i], rand)
        bs, K, H, N = a.shape
        h = w = int(round(N ** 0.5))
        a = a.contiguous().view(bs * K, H, h, w)
        a = R2.separable_conv2d(P, "%s.attn_convs0.%d" % (prefix, i), a, training)
        fm = F.conv2d(a, P["%s.attn_convs1.%d.weight" % (prefix, i)], P["%s.attn_convs1.%d.bias" % (prefix, i)])
        feat_maps.append(fm.reshape(bs, K, h, w))
        wq = P["%s.query_convs.%d.weight" % (prefix, i)]
        next_Q = F.linear(q, wq[:, :, 0], P["%s.query_convs.%d.bias" % (prefix, i)])
        updated_Qs.append(q.mean(dim=0, keepdim=True))
    return feat_maps, updated_Qs


def swin_unet_forward(P, x_lab, x_unlab=None, inference=False, cfg=SWIN_TINY_LITE, rand=None, training=True):
    """SwinUnet.forward (vision_transformer.py:90-108): 1 -> 3 channel repeat, two backbone passes, sspa(lab), sspa(unlab),
    uscl(unlab, queries of the labeled pass)."""
    rand = rand or R.NoRand()
    rep = (lambda t: t.repeat(1, 3, 1, 1) if t.shape[1] == 1 else t)
    out_lab, feats_lab = swin_sys_branch(P, rep(x_lab), cfg, rand, "swin_unet.")
    if inference:
        return out_lab
    out_unlab, feats_unlab = swin_sys_branch(P, rep(x_unlab), cfg, rand, "swin_unet.")
    maps_lab, Qs_lab = inherent_consistent_tokens(P, "sspa", feats_lab, None, "labeled", rand=rand, training=training)
    maps_consis, _ = inherent_consistent_tokens(P, "sspa", feats_unlab, None, "labeled", rand=rand, training=training)
    maps_unlab, _ = inherent_consistent_tokens(P, "uscl", feats_unlab, Qs_lab, "unlabeled", rand=rand, training=training)
    return out_lab, out_unlab, maps_lab, maps_unlab, maps_consis
