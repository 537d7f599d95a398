"""Swin-UNet backbone of the ICL path — drop-in for the reference's networks/swinunet_icl.py:
SwinTransformerSys (constructor :609, forward :796-809) and the blocks it is built from.

The module tree, parameter / buffer names and their order are the reference's (checkpoints interchange, including the
`relative_position_index` and `attn_mask` buffers).  Activations stay token-major [B, H*W, C] fp32 end to end:
  * LayerNorm, Linear (+ fused GELU) and the DropPath residuals are the icl_b200.functional kernels of the ICL heads;
  * window attention is ONE kernel per block (csrc/window_attn.cu) that folds the cyclic shift, the window partition /
    reverse and the shift mask into its addressing, so the reference's roll / view / permute copies do not exist;
  * patch merging / expanding are index permutations of the token tensor (torch views + one copy) around those kernels.
"""
import torch
import torch.nn as nn

from .. import functional as Fn
from .unet_3D_icl import MLP as Mlp  # fc1 -> GELU -> fc2, same parameter names (swinunet_icl.py:14-31)
from .unet_3D_icl import DropPath


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def _ln(t, m):
    return Fn.layer_norm(t, m.weight, m.bias, m.eps)


def _region_ids(n, ws, shift):
    """Wrap-around region (0, 1, 2) of every coordinate of the rolled frame: [0, n-ws), [n-ws, n-shift), [n-shift, n)."""
    c = torch.arange(n)
    return (c >= n - ws).long() + (c >= n - shift).long()


class WindowAttention(nn.Module):
    """Window multi-head self-attention with relative position bias (swinunet_icl.py:61-155)."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        if qk_scale or attn_drop or proj_drop:
            raise NotImplementedError("icl_b200 WindowAttention implements the reference's configuration (default qk scale, no dropout)")
        if window_size[0] != window_size[1]:
            raise NotImplementedError("icl_b200 WindowAttention: square windows only")
        self.dim, self.window_size, self.num_heads = dim, window_size, num_heads
        self.scale = (dim // num_heads) ** -0.5
        ws = window_size[0]
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) * (2 * ws - 1), num_heads))
        t = torch.arange(ws * ws)
        ty, tx = t // ws, t % ws
        index = (ty[:, None] - ty[None, :] + ws - 1) * (2 * ws - 1) + (tx[:, None] - tx[None, :] + ws - 1)
        self.register_buffer("relative_position_index", index)
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)
        self.softmax = nn.Softmax(dim=-1)

    def forward_tokens(self, x, H, W, shift):
        """x: [B, H*W, C] tokens at their image positions; the windows (rolled by `shift`) are formed inside the kernel."""
        qkv = Fn.linear(x, self.qkv.weight, self.qkv.bias)
        y = Fn.window_attention(qkv, self.relative_position_bias_table, H, W, self.num_heads, self.window_size[0], shift)
        return Fn.linear(y, self.proj.weight, self.proj.bias)

    def forward(self, x, mask=None):
        """The reference signature: x (num_windows*B, N, C) already partitioned.  Each window is attended as its own ws x ws image."""
        if mask is not None:
            raise NotImplementedError("icl_b200 WindowAttention.forward: masks are generated inside the kernel; call forward_tokens "
                                      "with the shift instead")
        ws = self.window_size[0]
        return self.forward_tokens(x, ws, ws, 0)


class SwinTransformerBlock(nn.Module):
    """swinunet_icl.py:183-293."""

    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop=0.0,
                 attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim, self.input_resolution, self.num_heads = dim, input_resolution, num_heads
        self.window_size, self.shift_size, self.mlp_ratio = window_size, shift_size, mlp_ratio
        if min(self.input_resolution) <= self.window_size:  # one window covers the map: no partition, no shift
            self.shift_size = 0
            self.window_size = min(self.input_resolution)
        assert 0 <= self.shift_size < self.window_size, "shift_size must in 0-window_size"
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, window_size=to_2tuple(self.window_size), num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale,
                                    attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        if self.shift_size > 0:
            # state_dict parity only (the kernel evaluates the mask from coordinates): -100 between tokens of one window
            # that come from different wrap-around regions of the rolled map (:217-245)
            H, W = self.input_resolution
            ws = self.window_size
            rid = _region_ids(H, ws, self.shift_size)[:, None] * 3 + _region_ids(W, ws, self.shift_size)[None, :]
            rid = rid.view(H // ws, ws, W // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
            attn_mask = torch.where(rid[:, None, :] != rid[:, :, None], -100.0, 0.0)
        else:
            attn_mask = None
        self.register_buffer("attn_mask", attn_mask)

    def _r(self, B, device):
        return self.drop_path.sample(B, device) if isinstance(self.drop_path, DropPath) else None

    def forward(self, x):
        H, W = self.input_resolution
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        y = self.attn.forward_tokens(_ln(x, self.norm1), H, W, self.shift_size)
        x = Fn.add_scaled(x, y, self._r(B, x.device))
        return Fn.add_scaled(x, self.mlp(_ln(x, self.norm2)), self._r(B, x.device))


class PatchMerging(nn.Module):
    """2x2 neighbourhood -> channels (order x[0::2,0::2], x[1::2,0::2], x[0::2,1::2], x[1::2,1::2]) -> LN(4C) -> Linear(4C, 2C)
    (swinunet_icl.py:310-351)."""

    def __init__(self, input_resolution, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.input_resolution, self.dim = input_resolution, dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(4 * dim)

    def forward(self, x):
        H, W = self.input_resolution
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        assert H % 2 == 0 and W % 2 == 0, f"x size ({H}*{W}) are not even."
        x = x.view(B, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 4, 2, 5).reshape(B, L // 4, 4 * C)  # channel block = dx * 2 + dy
        return Fn.linear(_ln(x, self.norm), self.reduction.weight, None)


def _pixel_shuffle_tokens(x, H, W, p):
    """'b h w (p1 p2 c) -> b (h p1) (w p2) c' on a token tensor [B, H*W, p*p*c]."""
    B, L, C = x.shape
    c = C // (p * p)
    return x.view(B, H, W, p, p, c).permute(0, 1, 3, 2, 4, 5).reshape(B, L * p * p, c)


class PatchExpand(nn.Module):
    """Linear(C, 2C) -> 2x2 pixel shuffle of the tokens -> LN(C/2)  (swinunet_icl.py:362-387)."""

    def __init__(self, input_resolution, dim, dim_scale=2, norm_layer=nn.LayerNorm):
        super().__init__()
        self.input_resolution, self.dim = input_resolution, dim
        self.expand = nn.Linear(dim, 2 * dim, bias=False) if dim_scale == 2 else nn.Identity()
        self.norm = norm_layer(dim // dim_scale)

    def forward(self, x):
        H, W = self.input_resolution
        if isinstance(self.expand, nn.Linear):
            x = Fn.linear(x, self.expand.weight, None)
        assert x.shape[1] == H * W, "input feature has wrong size"
        return _ln(_pixel_shuffle_tokens(x, H, W, 2), self.norm)


class FinalPatchExpand_X4(nn.Module):
    """Linear(C, 16C) -> 4x4 pixel shuffle -> LN(C)  (swinunet_icl.py:390-415)."""

    def __init__(self, input_resolution, dim, dim_scale=4, norm_layer=nn.LayerNorm):
        super().__init__()
        self.input_resolution, self.dim, self.dim_scale = input_resolution, dim, dim_scale
        self.expand = nn.Linear(dim, 16 * dim, bias=False)
        self.output_dim = dim
        self.norm = norm_layer(self.output_dim)

    def forward(self, x):
        H, W = self.input_resolution
        x = Fn.linear(x, self.expand.weight, None)
        assert x.shape[1] == H * W, "input feature has wrong size"
        return _ln(_pixel_shuffle_tokens(x, H, W, self.dim_scale), self.norm)


def _blocks(dim, input_resolution, depth, num_heads, window_size, mlp_ratio, qkv_bias, qk_scale, drop, attn_drop, drop_path, norm_layer):
    return nn.ModuleList([
        SwinTransformerBlock(dim=dim, input_resolution=input_resolution, num_heads=num_heads, window_size=window_size,
                             shift_size=0 if (i % 2 == 0) else window_size // 2, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                             drop=drop, attn_drop=attn_drop, drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                             norm_layer=norm_layer)
        for i in range(depth)])


class BasicLayer(nn.Module):
    """One encoder stage: `depth` Swin blocks (alternating shift 0 / ws//2) then optional PatchMerging (swinunet_icl.py:418-483)."""

    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop=0.0,
                 attn_drop=0.0, drop_path=0.0, norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False):
        super().__init__()
        self.dim, self.input_resolution, self.depth, self.use_checkpoint = dim, input_resolution, depth, use_checkpoint
        self.blocks = _blocks(dim, input_resolution, depth, num_heads, window_size, mlp_ratio, qkv_bias, qk_scale, drop, attn_drop, drop_path,
                              norm_layer)
        self.downsample = downsample(input_resolution, dim=dim, norm_layer=norm_layer) if downsample is not None else None

    def forward(self, x):
        for blk in self.blocks:
            x = blk(x)
        return self.downsample(x) if self.downsample is not None else x


class BasicLayer_up(nn.Module):
    """One decoder stage; also returns the block output before the PatchExpand (`inter_feat`, the ICL-head input)
    (swinunet_icl.py:486-553)."""

    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop=0.0,
                 attn_drop=0.0, drop_path=0.0, norm_layer=nn.LayerNorm, upsample=None, use_checkpoint=False):
        super().__init__()
        self.dim, self.input_resolution, self.depth, self.use_checkpoint = dim, input_resolution, depth, use_checkpoint
        self.blocks = _blocks(dim, input_resolution, depth, num_heads, window_size, mlp_ratio, qkv_bias, qk_scale, drop, attn_drop, drop_path,
                              norm_layer)
        self.upsample = PatchExpand(input_resolution, dim=dim, dim_scale=2, norm_layer=norm_layer) if upsample is not None else None

    def forward(self, x):
        for blk in self.blocks:
            x = blk(x)
        inter_feat = x
        if self.upsample is not None:
            x = self.upsample(x)
        return x, inter_feat


class PatchEmbed(nn.Module):
    """Conv2d(in_chans, embed_dim, kernel = stride = patch) + LayerNorm (swinunet_icl.py:556-603): non-overlapping patches, so the
    convolution is a Linear over the (c, ky, kx)-flattened patch."""

    def __init__(self, img_size=224, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        img_size, patch_size = to_2tuple(img_size), to_2tuple(patch_size)
        self.img_size, self.patch_size = img_size, patch_size
        self.patches_resolution = [img_size[0] // patch_size[0], img_size[1] // patch_size[1]]
        self.num_patches = self.patches_resolution[0] * self.patches_resolution[1]
        self.in_chans, self.embed_dim = in_chans, embed_dim
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None

    def forward(self, x):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        ph, pw = self.patch_size
        Hp, Wp = self.patches_resolution
        patches = x.view(B, C, Hp, ph, Wp, pw).permute(0, 2, 4, 1, 3, 5).reshape(B, Hp * Wp, C * ph * pw)
        x = Fn.linear(patches, self.proj.weight.reshape(self.embed_dim, C * ph * pw), self.proj.bias)
        return _ln(x, self.norm) if self.norm is not None else x


class SwinTransformerSys(nn.Module):
    """Swin-UNet (encoder, bottleneck, skip-connected expanding decoder, x4 head); swinunet_icl.py:606-809."""

    def __init__(self, img_size=224, patch_size=4, in_chans=3, num_classes=1000, embed_dim=96, depths=[2, 2, 2, 2],
                 depths_decoder=[1, 2, 2, 2], num_heads=[3, 6, 12, 24], window_size=7, mlp_ratio=4.0, qkv_bias=True, qk_scale=None,
                 drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.1, norm_layer=nn.LayerNorm, ape=False, patch_norm=True,
                 use_checkpoint=False, final_upsample="expand_first", **kwargs):
        super().__init__()
        if drop_rate or attn_drop_rate or use_checkpoint or final_upsample != "expand_first":
            raise NotImplementedError("icl_b200 SwinTransformerSys implements the reference configuration: drop_rate = attn_drop_rate = 0, "
                                      "no activation checkpointing, final_upsample='expand_first'")
        self.num_classes, self.num_layers, self.embed_dim = num_classes, len(depths), embed_dim
        self.ape, self.patch_norm = ape, patch_norm
        nl = self.num_layers
        self.num_features = int(embed_dim * 2 ** (nl - 1))
        self.num_features_up = int(embed_dim * 2)
        self.mlp_ratio, self.final_upsample = mlp_ratio, final_upsample
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                      norm_layer=norm_layer if self.patch_norm else None)
        num_patches = self.patch_embed.num_patches
        pr = self.patch_embed.patches_resolution
        self.patches_resolution = pr
        if self.ape:
            self.absolute_pos_embed = nn.Parameter(torch.zeros(1, num_patches, embed_dim))
            nn.init.trunc_normal_(self.absolute_pos_embed, std=0.02)
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]  # stochastic depth decay rule
        common = dict(window_size=window_size, mlp_ratio=self.mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate,
                      attn_drop=attn_drop_rate, norm_layer=norm_layer, use_checkpoint=use_checkpoint)
        self.layers = nn.ModuleList()
        for i in range(nl):
            self.layers.append(BasicLayer(dim=int(embed_dim * 2 ** i), input_resolution=(pr[0] // (2 ** i), pr[1] // (2 ** i)),
                                          depth=depths[i], num_heads=num_heads[i], drop_path=dpr[sum(depths[:i]):sum(depths[:i + 1])],
                                          downsample=PatchMerging if (i < nl - 1) else None, **common))
        self.layers_up = nn.ModuleList()
        self.concat_back_dim = nn.ModuleList()
        for i in range(nl):
            j = nl - 1 - i  # mirrored encoder stage
            dim, res = int(embed_dim * 2 ** j), (pr[0] // (2 ** j), pr[1] // (2 ** j))
            self.concat_back_dim.append(nn.Linear(2 * dim, dim) if i > 0 else nn.Identity())
            if i == 0:
                self.layers_up.append(PatchExpand(input_resolution=res, dim=dim, dim_scale=2, norm_layer=norm_layer))
            else:
                self.layers_up.append(BasicLayer_up(dim=dim, input_resolution=res, depth=depths[j], num_heads=num_heads[j],
                                                    drop_path=dpr[sum(depths[:j]):sum(depths[:j + 1])],
                                                    upsample=PatchExpand if (i < nl - 1) else None, **common))
        self.norm = norm_layer(self.num_features)
        self.norm_up = norm_layer(self.embed_dim)
        self.up = FinalPatchExpand_X4(input_resolution=(img_size // patch_size, img_size // patch_size), dim_scale=4, dim=embed_dim)
        self.output = nn.Conv2d(in_channels=embed_dim, out_channels=self.num_classes, kernel_size=1, bias=False)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"absolute_pos_embed"}

    @torch.jit.ignore
    def no_weight_decay_keywords(self):
        return {"relative_position_bias_table"}

    def forward_features(self, x):
        """Encoder + bottleneck (:752-765): returns the normalised bottleneck tokens and the stage inputs (skips)."""
        x = self.patch_embed(x)
        if self.ape:
            x = x + self.absolute_pos_embed
        x_downsample = []
        for layer in self.layers:
            x_downsample.append(x)
            x = layer(x)
        return _ln(x, self.norm), x_downsample

    def forward_up_features(self, x, x_downsample):
        """Decoder (:768-781): PatchExpand, then per stage cat([x, skip]) -> Linear -> blocks (-> PatchExpand)."""
        feats = []
        for inx, layer_up in enumerate(self.layers_up):
            if inx == 0:
                x = layer_up(x)
            else:
                cb = self.concat_back_dim[inx]
                x = Fn.linear(torch.cat([x, x_downsample[self.num_layers - 1 - inx]], -1), cb.weight, cb.bias)
                x, feat = layer_up(x)
                feats.append(feat)
        return _ln(x, self.norm_up), feats

    def up_x4(self, x):
        """x4 token expansion + 1x1 output convolution (:783-794); logits come back as an NCHW view of channels-last memory."""
        H, W = self.patches_resolution
        B, L, C = x.shape
        assert L == H * W, "input features has wrong size"
        x = self.up(x)
        y = Fn.linear(x, self.output.weight.reshape(self.num_classes, C), None)
        return y.view(B, 4 * H, 4 * W, self.num_classes).permute(0, 3, 1, 2)

    def _branch(self, x):
        t, skips = self.forward_features(x)
        last, feats = self.forward_up_features(t, skips)
        return self.up_x4(last), feats

    def forward(self, x_lab, x_unlab=None, inference=False):
        output_lab, feats_lab = self._branch(x_lab)
        if inference:
            return output_lab
        output_unlab, feats_unlab = self._branch(x_unlab)
        return output_lab, output_unlab, feats_lab, feats_unlab
