"""Forward/backward orchestration of the 3D U-Net backbone as ONE autograd node.

Mirrors the body shared by unet_3D.forward (reference networks/unet_3D.py:71-94) and each branch of
unet_3D_icl.forward (networks/unet_3D_icl.py:100-117): 9 UnetConv3 blocks (Conv3d 3^3 + InstanceNorm3d +
ReLU, twice; networks/utils.py:99-123), 4 MaxPool3d(2), 4 UnetUp3_CT (trilinear x2 + cat + UnetConv3;
networks/utils.py:260-276), Dropout(0.3) after `center` and `up1`, and the 1x1x1 `final` conv.

Why one node: all intermediate activations stay in the kernels' own formats (NDHWC fp32 + split-bf16 PK
operands), torch.cat / the upsampled tensor's concat are never materialised, and the hand-written backward
reproduces the reference's autograd pruning exactly — a branch whose `final` logits receive no gradient
(the unlabeled pass) skips the up_concat2/up_concat1/final backward and returns None for
their parameters, so `.grad is None`-ness (and therefore SGD weight-decay behaviour) matches.
"""
import torch

from .. import ops

PARAM_BLOCKS = ["conv1", "conv2", "conv3", "conv4", "center", "up_concat4.conv", "up_concat3.conv", "up_concat2.conv",
                "up_concat1.conv"]


def param_names():
    names = []
    for b in PARAM_BLOCKS:
        for h in ("conv1", "conv2"):
            names += ["%s.%s.0.weight" % (b, h), "%s.%s.0.bias" % (b, h)]
    names += ["final.weight", "final.bias"]
    return names


class _Act:
    """An activation in the kernel formats: f32 = NDHWC fp32 (None when only tensor-core consumers read it),
    pk = split-bf16 operand (None when C is not a multiple of 16)."""
    __slots__ = ("f32", "pk", "C", "shape")

    def __init__(self, f32, pk, shape=None):
        self.f32, self.pk = f32, pk
        self.shape = tuple(f32.shape) if f32 is not None else tuple(shape)
        self.C = self.shape[-1]


class _Half:
    __slots__ = ("srcs", "w", "y", "mr", "out", "umma", "owner")


def _half_fwd(srcs, w, b, owner=None):
    """conv3x3x3(+bias) -> InstanceNorm -> ReLU on the virtual concat of `srcs` (list of _Act).  `owner`: the nn.Parameter `w` was
    detached from (guards the packed-operand cache)."""
    cins = [s.C for s in srcs]
    cout = w.shape[0]
    B, D, H, W, _ = srcs[0].shape
    stats = ops.zeros((B, cout, 2), torch.float64, w.device)
    h = _Half()
    h.srcs, h.w, h.owner = srcs, w, owner
    h.umma = ops.umma_ok(cins, cout) and all(s.pk is not None for s in srcs)
    if h.umma:
        h.y = ops.conv3d_umma([s.pk for s in srcs], cins, ops.pack_w_umma_cached(w, False, D, owner), b, cout, B, D, H, W, stats)
    elif ops.stem_ok(cins, cout):
        h.y = ops.conv3d_stem_fwd(srcs[0].f32, w, b, B, D, H, W, stats)
    else:
        h.y = ops.conv3d_direct([s.f32 for s in srcs], cins, ops.repack_w_f32(w, False), b, cout, B, D, H, W, stats)
    h.mr = ops.instnorm_finalize(stats, B, cout, D * H * W)
    a, pk = ops.instnorm_relu_fwd(h.y, h.mr, ops.pk_ok(cout))
    h.out = _Act(a, pk)
    return h


def _half_bwd(h, dA, need_dx):
    """Returns (dw, db, [dx per source] or None)."""
    cins = [s.C for s in h.srcs]
    cout = h.w.shape[0]
    B, D, H, W, _ = h.y.shape
    dgrad_umma = need_dx and ops.umma_ok([cout], sum(cins), with_stats=False) and all(c % 16 == 0 for c in cins)
    wgrad_umma = ops.wgrad_umma_ok(cins, cout, D) and all(s.pk is not None for s in h.srcs)
    if wgrad_umma:
        # fp32 dY is only read by the CUDA-core data-gradient fallback
        dY, dY_pk, db = ops.instnorm_relu_bwd(dA, h.y, h.mr, True, want_dbias=True, want_f32=need_dx and not dgrad_umma)
        dw = ops.conv3d_wgrad_umma([s.pk for s in h.srcs], cins, dY_pk, cout, B, D, H, W)
    elif ops.stem_ok(cins, cout):
        dY, dY_pk, db = ops.instnorm_relu_bwd(dA, h.y, h.mr, dgrad_umma, want_dbias=True)
        dw = ops.conv3d_stem_wgrad(h.srcs[0].f32, dY, B, D, H, W)
    else:
        dY, dY_pk = ops.instnorm_relu_bwd(dA, h.y, h.mr, dgrad_umma)
        dw, db = ops.conv3d_wgrad([s.f32 for s in h.srcs], cins, dY, cout, B, D, H, W)
    if not need_dx:
        return dw, db, None
    cin_total = sum(cins)
    if dgrad_umma:
        wp = ops.pack_w_umma_cached(h.w, True, D, h.owner)
        if len(cins) == 2:
            d0, d1 = ops.conv3d_umma([dY_pk], [cout], wp, None, cin_total, B, D, H, W, split=cins[0])
            return dw, db, [d0, d1]
        return dw, db, [ops.conv3d_umma([dY_pk], [cout], wp, None, cin_total, B, D, H, W)]
    dx = ops.conv3d_direct([dY], [cout], ops.repack_w_f32(h.w, True), None, cin_total, B, D, H, W)
    if len(cins) == 2:
        return dw, db, [dx[..., :cins[0]].contiguous(), dx[..., cins[0]:].contiguous()]
    return dw, db, [dx]


def _block_fwd(srcs, p, owners=(None, None)):
    h1 = _half_fwd(srcs, p[0], p[1], owners[0])
    h2 = _half_fwd([h1.out], p[2], p[3], owners[1])
    return (h1, h2)


def _block_bwd(blk, dA, need_dx):
    h1, h2 = blk
    dw2, db2, dx2 = _half_bwd(h2, dA, True)
    dw1, db1, dx1 = _half_bwd(h1, dx2[0], need_dx)
    return [dw1, db1, dw2, db2], dx1


def _add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    ops.axpby(b, a, 1.0, 1.0)
    return a


class Backbone3DFn(torch.autograd.Function):
    """(x, drop_cfg, *38 params) -> (final [B,K,D,H,W], center_drop, up4, up3), all channels_last_3d views.

    drop_cfg: None (eval: Dropout is the identity) or (p, mask1, mask2, seed1, seed2) where maskN is a uint8
    keep-mask in NDHWC order or None (then Philox keyed by seedN draws it inside the kernel)."""

    @staticmethod
    def forward(ctx, x, drop_cfg, *params):
        ctx.set_materialize_grads(False)
        p = [t.detach() for t in params]
        blk = {name: p[4 * i:4 * i + 4] for i, name in enumerate(PARAM_BLOCKS)}
        own = {name: (params[4 * i], params[4 * i + 2]) for i, name in enumerate(PARAM_BLOCKS)}   # weight Parameters of conv1 / conv2
        wf, bf = p[36], p[37]
        x_ = ops.to_ndhwc(x.detach())
        B, D, H, W, Cin = x_.shape
        # every conv from the second one on runs on the tensor cores when all channel counts are multiples of 16:
        # pooled / upsampled tensors are then only ever read as PK operands and their fp32 copies are not written
        # (the tensor-core weight gradient needs an even depth at every level: with D % 32 != 0 the `center` level has an odd
        # depth, its weight gradient runs on the CUDA-core kernel and reads the fp32 copies, so they must exist)
        lean = all(ops.umma_ok([t.shape[1]], t.shape[0]) for t in p[2:36:2]) and ops.wgrad_umma_ok([16], 16, D // 16)
        if D % 16 or H % 16 or W % 16:
            raise RuntimeError("unet_3D backbone: spatial size must be a multiple of 16, got %s" % ((D, H, W),))
        rec = {}
        a0 = _Act(x_, ops.pack_pk(x_) if ops.pk_ok(Cin) else None)
        enc = a0
        for i, name in enumerate(["conv1", "conv2", "conv3", "conv4"]):
            rec[name] = _block_fwd([enc], blk[name], own[name])
            out = rec[name][1].out
            pooled, idx, ppk = ops.maxpool_fwd(out.f32, ops.pk_ok(out.C), want_f32=not lean)
            rec["pool%d" % (i + 1)] = idx
            enc = _Act(pooled, ppk, (B, out.shape[1] // 2, out.shape[2] // 2, out.shape[3] // 2, out.C))
        rec["center"] = _block_fwd([enc], blk["center"], own["center"])
        center = rec["center"][1].out
        if drop_cfg is not None:
            pdrop, m1, m2, s1, s2 = drop_cfg
            cd = ops.dropout(center.f32, pdrop, m1, s1)
            center_d = _Act(cd, ops.pack_pk(cd) if ops.pk_ok(center.C) else None)
        else:
            center_d = center
        coarse = center_d
        for name, skip in (("up_concat4.conv", "conv4"), ("up_concat3.conv", "conv3"), ("up_concat2.conv", "conv2"),
                           ("up_concat1.conv", "conv1")):
            up_f32, up_pk = ops.upsample2x_fwd(coarse.f32, ops.pk_ok(coarse.C), want_f32=not lean)
            cs = coarse.shape
            rec[name] = _block_fwd([rec[skip][1].out, _Act(up_f32, up_pk, (cs[0], 2 * cs[1], 2 * cs[2], 2 * cs[3], cs[4]))], blk[name], own[name])
            coarse = rec[name][1].out
        up1 = coarse
        up1d = ops.dropout(up1.f32, pdrop, m2, s2) if drop_cfg is not None else up1.f32
        K = wf.shape[0]
        rows = B * D * H * W
        final = torch.empty((B, D, H, W, K), dtype=torch.float32, device=x_.device)
        wf2 = wf.reshape(K, -1)
        if wf2.shape[1] == 16 and K <= 16:
            ops.call("icl_head1x1_fwd", ops.P(up1d), ops.P(wf2), ops.P(bf), ops.P(final), ops.c_ll(rows), ops.c_int(16), ops.c_int(K),
                     mbytes=4e-6 * rows * (16 + K))
        else:
            ops.sgemm(rows, K, wf2.shape[1], up1d, wf2.shape[1], 1, wf2, 1, wf2.shape[1], final, K, 1, bias=bf, bias_mode=1)
        rec["up1d"] = up1d
        ctx.rec, ctx.blk, ctx.wf, ctx.drop_cfg = rec, blk, wf2, drop_cfg
        ctx.needs_x = x.requires_grad
        outs = (final, center_d.f32, rec["up_concat4.conv"][1].out.f32, rec["up_concat3.conv"][1].out.f32)
        return tuple(ops.to_ncdhw_view(o) for o in outs)

    @staticmethod
    def backward(ctx, g_final, g_center, g_up4, g_up3):
        rec, wf2, drop_cfg = ctx.rec, ctx.wf, ctx.drop_cfg
        grads = {}
        nd = lambda g: None if g is None else ops.to_ndhwc(g)
        g_center, g_up4, g_up3 = nd(g_center), nd(g_up4), nd(g_up3)
        # gradients that reach a tensor from outside must not be modified in place
        own = lambda g: None if g is None else g.clone()
        d_c = {"conv1": None, "conv2": None, "conv3": None, "conv4": None}
        d_up3_in = d_up4_in = d_cd_in = None

        def up_block_bwd(name, skip, dA, coarse_shape):
            pg, dxs = _block_bwd(rec[name], dA, True)
            grads[name] = pg
            d_c[skip] = _add(d_c[skip], dxs[0])
            Bc, dc, hc, wc, Cc = coarse_shape
            dcoarse = torch.empty(coarse_shape, dtype=torch.float32, device=dA.device)
            ops.upsample2x_bwd(dxs[1], 0, Cc, dcoarse, False)
            return dcoarse

        if g_final is not None:
            gf = ops.to_ndhwc(g_final)
            B, D, H, W, K = gf.shape
            rows, C1 = B * D * H * W, wf2.shape[1]
            d_up1d = torch.empty((B, D, H, W, C1), dtype=torch.float32, device=gf.device)
            if C1 == 16 and K in (1, 2, 4, 16):
                dwf = ops.zeros((K, C1), torch.float32, gf.device)
                dbf = ops.zeros((K,), torch.float32, gf.device)
                ops.call("icl_head1x1_bwd", ops.P(gf), ops.P(rec["up1d"]), ops.P(wf2), ops.P(d_up1d), ops.P(dwf), ops.P(dbf), ops.c_ll(rows),
                         ops.c_int(16), ops.c_int(K), mbytes=4e-6 * rows * (32 + K))
            else:
                ops.sgemm(rows, C1, K, gf, K, 1, wf2, C1, 1, d_up1d, C1, 1)
                dwf = torch.empty((K, C1), dtype=torch.float32, device=gf.device)
                ops.sgemm(K, C1, rows, gf, 1, K, rec["up1d"], C1, 1, dwf, C1, 1)
                dbf = torch.empty((K,), dtype=torch.float32, device=gf.device)
                ops.call("icl_colsum", ops.P(gf), ops.P(dbf), ops.c_ll(rows), ops.c_int(K), ops.c_int(0), tag="%dx%d" % (rows, K))
            grads["final"] = [dwf.reshape(K, C1, 1, 1, 1), dbf]
            if drop_cfg is not None:
                d_up1 = ops.dropout(d_up1d, drop_cfg[0], drop_cfg[2], drop_cfg[4])
            else:
                d_up1 = d_up1d
            d_up2 = up_block_bwd("up_concat1.conv", "conv1", d_up1, rec["up_concat2.conv"][1].out.f32.shape)
            d_up3_in = up_block_bwd("up_concat2.conv", "conv2", d_up2, rec["up_concat3.conv"][1].out.f32.shape)
        d_up3 = _add(d_up3_in, g_up3) if d_up3_in is not None else own(g_up3)
        if d_up3 is not None:
            d_up4_in = up_block_bwd("up_concat3.conv", "conv3", d_up3, rec["up_concat4.conv"][1].out.f32.shape)
        d_up4 = _add(d_up4_in, g_up4) if d_up4_in is not None else own(g_up4)
        if d_up4 is not None:
            d_cd_in = up_block_bwd("up_concat4.conv", "conv4", d_up4, rec["center"][1].out.f32.shape)
        d_cd = _add(d_cd_in, g_center) if d_cd_in is not None else own(g_center)
        if d_cd is not None:
            d_center = ops.dropout(d_cd, drop_cfg[0], drop_cfg[1], drop_cfg[3]) if drop_cfg is not None else d_cd
            pg, dxs = _block_bwd(rec["center"], d_center, True)
            grads["center"] = pg
            d_pool = dxs[0]
            for i, name in reversed(list(enumerate(["conv1", "conv2", "conv3", "conv4"]))):
                out_shape = rec[name][1].out.f32.shape
                if d_c[name] is None:
                    d_c[name] = torch.empty(out_shape, dtype=torch.float32, device=d_pool.device)
                    ops.maxpool_bwd(d_pool, rec["pool%d" % (i + 1)], d_c[name], False)
                else:
                    ops.maxpool_bwd(d_pool, rec["pool%d" % (i + 1)], d_c[name], True)
                need_dx = (i > 0) or ctx.needs_x
                pg, dxs = _block_bwd(rec[name], d_c[name], need_dx)
                grads[name] = pg
                d_pool = dxs[0] if dxs is not None else None
            dx_in = ops.to_ncdhw_view(d_pool) if (ctx.needs_x and d_pool is not None) else None
        else:
            dx_in = None
        out = []
        for name in PARAM_BLOCKS:
            out += grads.get(name, [None, None, None, None])
        out += grads.get("final", [None, None])
        ctx.rec = None
        return (dx_in, None) + tuple(out)
