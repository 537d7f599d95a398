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
import os

import torch

from .. import lanes as Ln
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
    __slots__ = ("srcs", "w", "y", "mr", "out", "umma", "owner", "pk_bs")


def _half_fwd(srcs, w, b, owner=None, want_f32=True):
    """conv3x3x3(+bias) -> InstanceNorm -> ReLU on the virtual concat of `srcs` (list of _Act).  `owner`: the nn.Parameter `w` was
    detached from (guards the packed-operand cache).  want_f32=False: the output is only ever read as a PK operand."""
    cins = [s.C for s in srcs]
    cout = w.shape[0]
    B, D, H, W, _ = srcs[0].shape
    stats = ops.zeros((B, cout, 2), torch.float64, w.device)
    h = _Half()
    h.srcs, h.w, h.owner, h.pk_bs = srcs, w, owner, 0
    h.umma = ops.umma_ok(cins, cout) and all(s.pk is not None for s in srcs)
    if h.umma:
        h.y = ops.conv3d_umma([s.pk for s in srcs], cins, ops.pack_w_umma_cached(w, False, D, owner), b, cout, B, D, H, W, stats)
    elif ops.stem_ok(cins, cout):
        h.y = ops.conv3d_stem_fwd(srcs[0].f32, w, b, B, D, H, W, stats)
    else:
        h.y = ops.conv3d_direct([s.f32 for s in srcs], cins, ops.repack_w_f32(w, False), b, cout, B, D, H, W, stats)
    h.mr = ops.instnorm_finalize(stats, B, cout, D * H * W)
    a, pk = ops.instnorm_relu_fwd(h.y, h.mr, ops.pk_ok(cout), want_f32=want_f32 or not ops.pk_ok(cout))
    h.out = _Act(a, pk, tuple(h.y.shape))
    return h


def _half_prefix(h, n):
    """The first n samples of a recorded conv half (batch-major fp32 tensors are contiguous prefixes; the PK operands stay the
    full tensors and are read with their allocated batch as plane stride, `pk_bs`)."""
    if n == h.y.shape[0]:
        return h
    q = _Half()
    q.srcs = [_Act(None if s.f32 is None else s.f32[:n], s.pk, (n,) + tuple(s.shape[1:])) for s in h.srcs]
    q.w, q.owner, q.umma, q.out = h.w, h.owner, h.umma, None
    q.y, q.mr = h.y[:n], h.mr[:n]
    q.pk_bs = h.y.shape[0]
    return q


def _wgrad_lane(device):
    if os.environ.get("ICL_WGRAD_LANE", "1") == "0" or not Ln.enabled(device):
        return None, None
    L = Ln.get(device, ["w"], priority=0)
    return L["w"], L["main"]


def _wgrad_join(device):
    wl, main = _wgrad_lane(device)
    Ln.handoff(main, wl)


def _half_bwd(h, dA, need_dx, dx_out0=None):
    """Returns (dw, db, [dx per source] or None).  dx_out0: optional preallocated tensor for the gradient of the FIRST source."""
    cins = [s.C for s in h.srcs]
    cout = h.w.shape[0]
    B, D, H, W, _ = h.y.shape
    dgrad_umma = need_dx and ops.umma_ok([cout], sum(cins), with_stats=False) and all(c % 16 == 0 for c in cins)
    wgrad_umma = ops.wgrad_umma_ok(cins, cout, D) and all(s.pk is not None for s in h.srcs)
    if wgrad_umma:
        # fp32 dY is only read by the CUDA-core data-gradient fallback
        dY, dY_pk, db = ops.instnorm_relu_bwd(dA, h.y, h.mr, True, want_dbias=True, want_f32=need_dx and not dgrad_umma)
        # the weight gradient is a leaf of the backward chain: it runs on a lane of its own next to the data gradient that continues
        # the chain (at the deep levels neither fills the GPU alone); the caller joins the lane before backward returns (_wgrad_join)
        wl, main = _wgrad_lane(dY_pk.device)
        Ln.handoff(wl, main, dY_pk)
        with Ln.on(wl):
            dw = ops.conv3d_wgrad_umma([s.pk for s in h.srcs], cins, dY_pk, cout, B, D, H, W, bx=h.pk_bs)
        if wl is not None:
            dw.record_stream(main)
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
            d0, d1 = ops.conv3d_umma([dY_pk], [cout], wp, None, cin_total, B, D, H, W, split=cins[0], out=dx_out0)
            return dw, db, [d0, d1]
        return dw, db, [ops.conv3d_umma([dY_pk], [cout], wp, None, cin_total, B, D, H, W, out=dx_out0)]
    dx = ops.conv3d_direct([dY], [cout], ops.repack_w_f32(h.w, True), None, cin_total, B, D, H, W)
    if len(cins) == 2:
        d0 = dx[..., :cins[0]].contiguous()
        if dx_out0 is not None:
            dx_out0.copy_(d0)
            d0 = dx_out0
        return dw, db, [d0, dx[..., cins[0]:].contiguous()]
    if dx_out0 is not None:
        dx_out0.copy_(dx)
        dx = dx_out0
    return dw, db, [dx]


def _block_fwd(srcs, p, owners=(None, None), lean=False):
    # lean: the second convolution (forward, data gradient and weight gradient) runs on the tensor cores, so the activation
    # between the two convolutions is never read in fp32
    h1 = _half_fwd(srcs, p[0], p[1], owners[0], want_f32=not lean)
    h2 = _half_fwd([h1.out], p[2], p[3], owners[1])
    return (h1, h2)


def _block_bwd(blk, dA, need_dx, n=None, dx_out0=None):
    """n: run on the first n samples only (labeled prefix of a batched labeled + unlabeled pass)."""
    h1, h2 = blk
    if n is not None:
        h1, h2 = _half_prefix(h1, n), _half_prefix(h2, n)
    dw2, db2, dx2 = _half_bwd(h2, dA, True)
    dw1, db1, dx1 = _half_bwd(h1, dx2[0], need_dx, dx_out0)
    return [dw1, db1, dw2, db2], dx1


def _add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    ops.axpby(b, a, 1.0, 1.0)
    return a


LOW_BLOCKS = PARAM_BLOCKS[:7]    # encoder, center, up_concat4, up_concat3: everything the ICL heads read
TOP_BLOCKS = PARAM_BLOCKS[7:]    # up_concat2, up_concat1 (+ final): 48^3 / 96^3 levels, never seen by the heads
N_LOW = 4 * len(LOW_BLOCKS)


def _forward_low(st, x, drop_cfg, params_low, top_weights):
    """Encoder, center (+ dropout1), up_concat4, up_concat3 on the batch x; records what backward and _forward_top need on `st`
    (an autograd ctx or a _PairState).  `top_weights`: the conv weights of up_concat2 / up_concat1 (only their shapes are read: the
    activation formats are chosen for the whole network).  Returns the NDHWC tensors (center after dropout1, up4, up3)."""
    p = [t.detach() for t in params_low]
    st.blk = {name: p[4 * i:4 * i + 4] for i, name in enumerate(LOW_BLOCKS)}
    st.own = {name: (params_low[4 * i], params_low[4 * i + 2]) for i, name in enumerate(LOW_BLOCKS)}   # weight Parameters of conv1 / conv2
    blk, own = st.blk, st.own
    x_ = ops.to_ndhwc(x.detach())
    B, D, H, W, Cin = x_.shape
    # every conv from the second one on runs on the tensor cores when all channel counts are multiples of 16:
    # pooled / upsampled tensors are then only ever read as PK operands and their fp32 copies are not written
    # (the tensor-core weight gradient needs an even depth at every level: with D % 32 != 0 the `center` level has an odd
    # depth, its weight gradient runs on the CUDA-core kernel and reads the fp32 copies, so they must exist)
    lean = all(ops.umma_ok([t.shape[1]], t.shape[0]) for t in p[2::2] + list(top_weights)) and ops.wgrad_umma_ok([16], 16, D // 16)
    if D % 16 or H % 16 or W % 16:
        raise RuntimeError("unet_3D backbone: spatial size must be a multiple of 16, got %s" % ((D, H, W),))
    rec = {}
    a0 = _Act(x_, ops.pack_pk(x_) if ops.pk_ok(Cin) else None)
    enc = a0
    for i, name in enumerate(["conv1", "conv2", "conv3", "conv4"]):
        rec[name] = _block_fwd([enc], blk[name], own[name], lean)
        out = rec[name][1].out
        pooled, idx, ppk = ops.maxpool_fwd(out.f32, ops.pk_ok(out.C), want_f32=not lean)
        rec["pool%d" % (i + 1)] = idx
        enc = _Act(pooled, ppk, (B, out.shape[1] // 2, out.shape[2] // 2, out.shape[3] // 2, out.C))
    rec["center"] = _block_fwd([enc], blk["center"], own["center"], lean)
    center = rec["center"][1].out
    if drop_cfg is not None:
        pdrop, m1, m2, s1, s2 = drop_cfg
        cd = ops.dropout(center.f32, pdrop, m1, s1)
        center_d = _Act(cd, ops.pack_pk(cd) if ops.pk_ok(center.C) else None)
    else:
        center_d = center
    coarse = center_d
    for name, skip in (("up_concat4.conv", "conv4"), ("up_concat3.conv", "conv3")):
        coarse = _up_fwd(rec, name, skip, coarse, blk, own, lean)
    st.rec, st.drop_cfg, st.lean = rec, drop_cfg, lean
    st.needs_x = x.requires_grad
    return (center_d.f32, rec["up_concat4.conv"][1].out.f32, rec["up_concat3.conv"][1].out.f32)


def _up_fwd(rec, name, skip, coarse, blk, own, lean):
    up_f32, up_pk = ops.upsample2x_fwd(coarse.f32, ops.pk_ok(coarse.C), want_f32=not lean)
    cs = coarse.shape
    rec[name] = _block_fwd([rec[skip][1].out, _Act(up_f32, up_pk, (cs[0], 2 * cs[1], 2 * cs[2], 2 * cs[3], cs[4]))], blk[name], own[name], lean)
    return rec[name][1].out


def _forward_top(st, params_top):
    """up_concat2, up_concat1, dropout2 and the 1x1x1 `final` conv on top of what _forward_low recorded.  Returns the NDHWC logits."""
    p = [t.detach() for t in params_top]
    for i, name in enumerate(TOP_BLOCKS):
        st.blk[name] = p[4 * i:4 * i + 4]
        st.own[name] = (params_top[4 * i], params_top[4 * i + 2])
    wf, bf = p[8], p[9]
    rec, drop_cfg = st.rec, st.drop_cfg
    coarse = rec["up_concat3.conv"][1].out
    for name, skip in (("up_concat2.conv", "conv2"), ("up_concat1.conv", "conv1")):
        coarse = _up_fwd(rec, name, skip, coarse, st.blk, st.own, st.lean)
    up1 = coarse
    up1d = ops.dropout(up1.f32, drop_cfg[0], drop_cfg[2], drop_cfg[4]) if drop_cfg is not None else up1.f32
    B, D, H, W, _ = up1.shape
    K = wf.shape[0]
    rows = B * D * H * W
    final = torch.empty((B, D, H, W, K), dtype=torch.float32, device=up1d.device)
    wf2 = wf.reshape(K, -1)
    if wf2.shape[1] == 16 and K <= 16:
        ops.call("icl_head1x1_fwd", ops.P(up1d), ops.P(wf2), ops.P(bf), ops.P(final), ops.c_ll(rows), ops.c_int(16), ops.c_int(K),
                 mbytes=4e-6 * rows * (16 + K))
    else:
        ops.sgemm(rows, K, wf2.shape[1], up1d, wf2.shape[1], 1, wf2, 1, wf2.shape[1], final, K, 1, bias=bf, bias_mode=1)
    rec["up1d"] = up1d
    st.wf = wf2
    return final


def _forward_body(ctx, x, drop_cfg, params):
    """Forward of the whole backbone on the batch x; records everything backward needs on ctx.  Returns the NDHWC tensors
    (final logits, center after dropout1, up4, up3)."""
    ctx.set_materialize_grads(False)
    low = _forward_low(ctx, x, drop_cfg, params[:N_LOW], params[N_LOW:N_LOW + 8:2])
    return (_forward_top(ctx, params[N_LOW:]),) + low


class Backbone3DFn(torch.autograd.Function):
    """(x, drop_cfg, *38 params) -> (final [B,K,D,H,W], center_drop, up4, up3), all channels_last_3d views.

    drop_cfg: None (eval: Dropout is the identity) or (p, mask1, mask2, seed1, seed2) where maskN is a uint8
    keep-mask in NDHWC order or None (then Philox keyed by seedN draws it inside the kernel)."""

    @staticmethod
    def forward(ctx, x, drop_cfg, *params):
        outs = _forward_body(ctx, x, drop_cfg, params)
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
        _wgrad_join(wf2.device)
        return (dx_in, None) + tuple(out)


def _head_bwd(gf, up1d, wf2):
    """Backward of the `final` 1x1x1 conv on NDHWC tensors: returns (d_up1d, [dW, db])."""
    B, D, H, W, K = gf.shape
    rows, C1 = B * D * H * W, wf2.shape[1]
    d_up1d = torch.empty((B, D, H, W, C1), dtype=torch.float32, device=gf.device)
    if C1 == 16 and K in (1, 2, 4, 16):
        dwf = ops.zeros((K, C1), torch.float32, gf.device)
        dbf = ops.zeros((K,), torch.float32, gf.device)
        ops.call("icl_head1x1_bwd", ops.P(gf), ops.P(up1d), ops.P(wf2), ops.P(d_up1d), ops.P(dwf), ops.P(dbf), ops.c_ll(rows),
                 ops.c_int(16), ops.c_int(K), mbytes=4e-6 * rows * (32 + K))
    else:
        ops.sgemm(rows, C1, K, gf, K, 1, wf2, C1, 1, d_up1d, C1, 1)
        dwf = torch.empty((K, C1), dtype=torch.float32, device=gf.device)
        ops.sgemm(K, C1, rows, gf, 1, K, up1d, C1, 1, dwf, C1, 1)
        dbf = torch.empty((K,), dtype=torch.float32, device=gf.device)
        ops.call("icl_colsum", ops.P(gf), ops.P(dbf), ops.c_ll(rows), ops.c_int(K), ops.c_int(0), tag="%dx%d" % (rows, K))
    return d_up1d, [dwf.reshape(K, C1, 1, 1, 1), dbf]


def _assemble(shape_full, parts, nl, dev):
    """Full-batch gradient from row-range pieces [(row0, row1, tensor)]; rows nobody covers are zero.  None if no piece.
    Pieces are cut at the labeled / unlabeled boundary, so every piece covers exactly one of the two segments."""
    cut = []
    for a, b_, t in parts:
        if t is None or a == b_:
            continue
        if a < nl < b_:
            cut += [(a, nl, t[:nl - a]), (nl, b_, t[nl - a:])]
        else:
            cut.append((a, b_, t))
    if not cut:
        return None
    out = torch.empty(shape_full, dtype=torch.float32, device=dev)
    for seg in ((0, nl), (nl, shape_full[0])):
        if seg[0] == seg[1]:
            continue
        mine = [t for a, b_, t in cut if (a, b_) == seg]
        if not mine:
            out[seg[0]:seg[1]].zero_()
            continue
        out[seg[0]:seg[1]].copy_(mine[0])
        for t in mine[1:]:
            ops.axpby(t.contiguous(), out[seg[0]:seg[1]], 1.0, 1.0)
    return out


_nd = lambda g: None if g is None else ops.to_ndhwc(g)


def _pair_bwd_top(st, nl, B, gfl, gfu):
    """Backward of final, dropout2, up_concat1, up_concat2 on the samples whose logits carry a gradient (the labeled prefix unless
    final_unlab received one).  Returns (parameter gradients {block: [...]}, nh, skip-gradient buffers, d_up3 of the first nh samples)."""
    rec, wf2, drop_cfg = st.rec, st.wf, st.drop_cfg
    dev = wf2.device
    grads = {}
    gfl, gfu = _nd(gfl), _nd(gfu)
    nh = 0
    if gfu is not None:
        nh = B
        gf = _assemble((B,) + tuple(gfu.shape[1:]), [(0, nl, gfl), (nl, B, gfu)], nl, dev)
    elif gfl is not None:
        nh, gf = nl, gfl
    d_skip = {"conv1": None, "conv2": None}   # full-batch buffers whose first nh samples hold the skip gradients
    d_up3_head = None
    if nh:
        d_up1d, grads["final"] = _head_bwd(gf, rec["up1d"][:nh], wf2)
        if drop_cfg is not None:
            m2 = None if drop_cfg[2] is None else drop_cfg[2][:nh]
            d_up = ops.dropout(d_up1d, drop_cfg[0], m2, drop_cfg[4])
        else:
            d_up = d_up1d
        for name, skip, below in (("up_concat1.conv", "conv1", "up_concat2.conv"), ("up_concat2.conv", "conv2", "up_concat3.conv")):
            full = torch.empty(rec[skip][1].out.shape, dtype=torch.float32, device=dev)
            pg, dxs = _block_bwd(rec[name], d_up, True, n=nh, dx_out0=full[:nh])
            grads[name] = pg
            d_skip[skip] = full
            cs = rec[below][1].out.shape
            d_up = torch.empty((nh,) + tuple(cs[1:]), dtype=torch.float32, device=dev)
            ops.upsample2x_bwd(dxs[1], 0, cs[4], d_up, False)
        d_up3_head = d_up
    return grads, nh, d_skip, d_up3_head


def _pair_bwd_low(st, nl, B, nh, d_skip, d_up3_head, gcl, gcu, g4l, g4u, g3l, g3u):
    """Backward of the layers both branches reach (up_concat3/4, center, encoder) on the whole batch, with the per-branch gradients
    assembled per row range.  Returns (parameter gradients {block: [...]}, dx or None)."""
    rec, drop_cfg = st.rec, st.drop_cfg
    dev = rec["center"][1].y.device
    grads = {}

    def up_block_bwd(name, dA, coarse_shape):
        pg, dxs = _block_bwd(rec[name], dA, True)
        grads[name] = pg
        dcoarse = torch.empty(coarse_shape, dtype=torch.float32, device=dev)
        ops.upsample2x_bwd(dxs[1], 0, coarse_shape[4], dcoarse, False)
        return dxs[0], dcoarse

    s3, s4, sc = rec["up_concat3.conv"][1].out.shape, rec["up_concat4.conv"][1].out.shape, rec["center"][1].out.shape
    d_c3 = d_c4 = None
    d_up3 = _assemble(tuple(s3), [(0, nl, _nd(g3l)), (nl, B, _nd(g3u)), (0, nh, d_up3_head)], nl, dev)
    d_up4_in = d_cd_in = None
    if d_up3 is not None:
        d_c3, d_up4_in = up_block_bwd("up_concat3.conv", d_up3, tuple(s4))
    d_up4 = _assemble(tuple(s4), [(0, nl, _nd(g4l)), (nl, B, _nd(g4u)), (0, B, d_up4_in)], nl, dev)
    if d_up4 is not None:
        d_c4, d_cd_in = up_block_bwd("up_concat4.conv", d_up4, tuple(sc))
    d_cd = _assemble(tuple(sc), [(0, nl, _nd(gcl)), (nl, B, _nd(gcu)), (0, B, d_cd_in)], nl, dev)
    dx_in = None
    if d_cd is not None:
        d_center = ops.dropout(d_cd, drop_cfg[0], drop_cfg[1], drop_cfg[3]) if drop_cfg is not None else d_cd
        pg, dxs = _block_bwd(rec["center"], d_center, True)
        grads["center"] = pg
        d_pool = dxs[0]
        d_full = {"conv4": d_c4, "conv3": d_c3, "conv2": d_skip["conv2"], "conv1": d_skip["conv1"]}
        for i, name in reversed(list(enumerate(["conv1", "conv2", "conv3", "conv4"]))):
            idx = rec["pool%d" % (i + 1)]
            dc = d_full[name]
            if dc is None:
                dc = torch.empty(rec[name][1].out.shape, dtype=torch.float32, device=dev)
                ops.maxpool_bwd(d_pool, idx, dc, False)
            elif name in ("conv1", "conv2") and nh < B:
                # skip gradients exist for the first nh samples only: accumulate there, plain write for the rest
                ops.maxpool_bwd(d_pool[:nh], idx[:nh], dc[:nh], True)
                ops.maxpool_bwd(d_pool[nh:], idx[nh:], dc[nh:], False)
            else:
                ops.maxpool_bwd(d_pool, idx, dc, True)
            need_dx = (i > 0) or st.needs_x
            pg, dxs = _block_bwd(rec[name], dc, need_dx)
            grads[name] = pg
            d_pool = dxs[0] if dxs is not None else None
        dx_in = ops.to_ncdhw_view(d_pool) if (st.needs_x and d_pool is not None) else None
    return grads, dx_in


def _flat(grads, names, with_final=False):
    out = []
    for name in names:
        out += grads.get(name, [None, None, None, None])
    if with_final:
        out += grads.get("final", [None, None])
    return out


class BackbonePairFn(torch.autograd.Function):
    """The labeled and the unlabeled pass of unet_3D_icl.forward (unet_3D_icl.py:100-139) as ONE batched backbone pass:
    (x = cat([x_lab, x_unlab]), n_lab, drop_cfg, *38 params) ->
    (final_lab, final_unlab, center_lab, center_unlab, up4_lab, up4_unlab, up3_lab, up3_unlab).

    Exact: InstanceNorm is per sample and both passes use the same weights, so batching changes no value — it halves the launches
    of the forward pass and doubles the tiles per convolution launch.  Backward reproduces the reference's pruning with sub-batches
    instead of two autograd nodes: `final`, up_concat1 and up_concat2 run on the samples whose logits received a gradient (the
    labeled prefix: final_unlab is only ever used detached), the layers both branches reach (up_concat3/4, center,
    encoder) run once on the whole batch with the per-branch gradients assembled per row range.  Parameter gradients are the sums
    over the samples that took part, i.e. exactly what autograd accumulates from the reference's two passes."""

    @staticmethod
    def forward(ctx, x, n_lab, drop_cfg, *params):
        final, center, up4, up3 = _forward_body(ctx, x, drop_cfg, params)
        ctx.n_lab, ctx.B = int(n_lab), final.shape[0]
        v = ops.to_ncdhw_view
        n = ctx.n_lab
        return (v(final[:n]), v(final[n:]), v(center[:n]), v(center[n:]), v(up4[:n]), v(up4[n:]), v(up3[:n]), v(up3[n:]))

    @staticmethod
    def backward(ctx, gfl, gfu, gcl, gcu, g4l, g4u, g3l, g3u):
        nl, B = ctx.n_lab, ctx.B
        gt, nh, d_skip, d_up3_head = _pair_bwd_top(ctx, nl, B, gfl, gfu)
        gl, dx_in = _pair_bwd_low(ctx, nl, B, nh, d_skip, d_up3_head, gcl, gcu, g4l, g4u, g3l, g3u)
        gl.update(gt)
        ctx.rec = None
        _wgrad_join(ctx.wf.device)
        return (dx_in, None, None) + tuple(_flat(gl, PARAM_BLOCKS, with_final=True))


class _PairState:
    """What BackbonePairLowFn and BackbonePairTopFn share: the recorded activations of the forward pass and, during backward, the
    gradients the top section hands down (skip connections of conv1 / conv2, d up3 of the samples that went through it)."""

    def __init__(self):
        self.rec = self.blk = self.own = self.wf = self.drop_cfg = None
        self.lean = self.needs_x = False
        self.n_lab = self.B = 0
        self.top = None   # (nh, d_skip, d_up3_head) once BackbonePairTopFn.backward has run


class BackbonePairLowFn(torch.autograd.Function):
    """BackbonePairFn cut in two autograd nodes, lower part: (x, n_lab, drop_cfg, state, top conv weights' shapes, *28 params of
    conv1..conv4, center, up_concat4, up_concat3) -> (center_lab, center_unlab, up4_lab, up4_unlab, up3_lab, up3_unlab, link).

    Why two nodes: the ICL heads read center / up4 / up3 only.  With the upper decoder levels (up_concat2, up_concat1, final: the
    48^3 and 96^3 layers, about half of the backbone's time) in a node of their own, the heads' ~650 small kernels run on side streams
    next to them in the forward pass, and the heads' backward runs next to the upper levels' backward (icl_b200/lanes.py).  Same
    kernels on the same inputs as the single node.  `link` is a one-element tensor whose only purpose is the autograd edge that makes
    the top node's backward run before this node's; the gradients the top section produces for this section travel on the state."""

    @staticmethod
    def forward(ctx, x, n_lab, drop_cfg, st, top_weights, *params_low):
        ctx.set_materialize_grads(False)
        center, up4, up3 = _forward_low(st, x, drop_cfg, params_low, top_weights)
        st.n_lab, st.B = int(n_lab), center.shape[0]
        ctx.st = st
        v = ops.to_ncdhw_view
        n = st.n_lab
        link = torch.empty((1,), dtype=torch.float32, device=center.device)
        return (v(center[:n]), v(center[n:]), v(up4[:n]), v(up4[n:]), v(up3[:n]), v(up3[n:]), link)

    @staticmethod
    def backward(ctx, gcl, gcu, g4l, g4u, g3l, g3u, glink):
        st = ctx.st
        nh, d_skip, d_up3_head = st.top if st.top is not None else (0, {"conv1": None, "conv2": None}, None)
        gl, dx_in = _pair_bwd_low(st, st.n_lab, st.B, nh, d_skip, d_up3_head, gcl, gcu, g4l, g4u, g3l, g3u)
        st.rec = st.top = None
        _wgrad_join(st.wf.device)
        return (dx_in, None, None, None, None) + tuple(_flat(gl, LOW_BLOCKS))


class BackbonePairTopFn(torch.autograd.Function):
    """Upper part: (link, state, *10 params of up_concat2, up_concat1, final) -> (final_lab, final_unlab)."""

    @staticmethod
    def forward(ctx, link, st, *params_top):
        ctx.set_materialize_grads(False)
        final = _forward_top(st, params_top)
        ctx.st = st
        v = ops.to_ncdhw_view
        n = st.n_lab
        return v(final[:n]), v(final[n:])

    @staticmethod
    def backward(ctx, gfl, gfu):
        st = ctx.st
        gt, nh, d_skip, d_up3_head = _pair_bwd_top(st, st.n_lab, st.B, gfl, gfu)
        st.top = (nh, d_skip, d_up3_head)
        _wgrad_join(st.wf.device)
        return (None, None) + tuple(_flat(gt, TOP_BLOCKS, with_final=True))
