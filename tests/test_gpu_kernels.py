"""GPU parity tests, kernel level: every C-ABI kernel against the torch-CPU op it replaces
(the arithmetic the reference dispatches), on seeded inputs at sizes the CPU finishes in seconds.
Tolerances: fp32 kernels 1e-4 rel-L2; tcgen05 convolution 2e-4 in "parity" (bf16x3) mode and the
stated loose bound 3e-2 in "fast" (single bf16) mode; integer outputs bit-exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import assert_close, rel_l2

pytestmark = pytest.mark.gpu


def _ops():
    from icl_b200 import ops
    return ops


def g(seed):
    return torch.Generator().manual_seed(seed)


def cl(x):  # NCDHW cpu -> NDHWC cuda
    return x.permute(0, 2, 3, 4, 1).contiguous().cuda()


def uncl(x):  # NDHWC cuda -> NCDHW cpu
    return x.permute(0, 4, 1, 2, 3).contiguous().cpu()


# ------------------------------------------------------------------------------------------ conv
@pytest.mark.parametrize("B,cins,cout,dims", [
    (2, [1], 16, (8, 12, 20)), (1, [5], 7, (5, 9, 17)), (1, [8, 12], 20, (6, 8, 16)), (2, [16], 16, (4, 16, 8)),
])
def test_conv3d_direct_fwd(B, cins, cout, dims):
    ops = _ops()
    D, H, W = dims
    xs = [torch.randn(B, c, D, H, W, generator=g(i)) for i, c in enumerate(cins)]
    w = torch.randn(cout, sum(cins), 3, 3, 3, generator=g(9)) * 0.1
    b = torch.randn(cout, generator=g(10))
    ref = F.conv3d(torch.cat(xs, 1), w, b, padding=1)
    stats = torch.zeros(B, cout, 2, dtype=torch.float64, device="cuda")
    y = ops.conv3d_direct([cl(x) for x in xs], cins, ops.repack_w_f32(w.cuda(), False), b.cuda(), cout, B, D, H, W, stats)
    assert_close(uncl(y), ref, 1e-5, "direct fwd")
    assert_close(stats[..., 0].cpu(), ref.double().sum((2, 3, 4)), 1e-5, "sum", abs_floor=1e-3)
    assert_close(stats[..., 1].cpu(), (ref.double() ** 2).sum((2, 3, 4)), 1e-5, "sumsq")


def test_conv3d_direct_dgrad_and_wgrad():
    ops = _ops()
    B, cin, cout, D, H, W = 2, 6, 10, 6, 10, 9
    x = torch.randn(B, cin, D, H, W, generator=g(1), requires_grad=True)
    w = (torch.randn(cout, cin, 3, 3, 3, generator=g(2)) * 0.1).requires_grad_(True)
    b = torch.randn(cout, generator=g(3)).requires_grad_(True)
    dy = torch.randn(B, cout, D, H, W, generator=g(4))
    F.conv3d(x, w, b, padding=1).backward(dy)
    dx = ops.conv3d_direct([cl(dy)], [cout], ops.repack_w_f32(w.detach().cuda(), True), None, cin, B, D, H, W)
    assert_close(uncl(dx), x.grad, 1e-5, "dgrad")
    dw, db = ops.conv3d_wgrad([cl(x.detach())], [cin], cl(dy), cout, B, D, H, W)
    assert_close(dw.cpu(), w.grad, 1e-5, "wgrad")
    assert_close(db.cpu(), b.grad, 1e-5, "bgrad")


@pytest.mark.parametrize("B,dims", [(2, (8, 16, 32)), (1, (6, 12, 40)), (1, (4, 9, 33)), (2, (32, 64, 96))])   # last: more tiles than CTAs
def test_conv3d_stem_fwd_and_wgrad(B, dims):
    """Cin=1 -> 16 stem specialisations vs F.conv3d and its autograd."""
    ops = _ops()
    D, H, W = dims
    x = torch.randn(B, 1, D, H, W, generator=g(1))
    w = (torch.randn(16, 1, 3, 3, 3, generator=g(2)) * 0.3).requires_grad_(True)
    b = torch.randn(16, generator=g(3))
    dy = torch.randn(B, 16, D, H, W, generator=g(4))
    ref = F.conv3d(x, w, b, padding=1)
    ref.backward(dy)
    stats = torch.zeros(B, 16, 2, dtype=torch.float64, device="cuda")
    y = ops.conv3d_stem_fwd(cl(x), w.detach().cuda(), b.cuda(), B, D, H, W, stats)
    assert_close(uncl(y), ref.detach(), 1e-5, "stem fwd")
    assert_close(stats[..., 0].cpu(), ref.detach().double().sum((2, 3, 4)), 1e-5, "sum", abs_floor=1e-3)
    assert_close(stats[..., 1].cpu(), (ref.detach().double() ** 2).sum((2, 3, 4)), 1e-5, "sumsq")
    dw = ops.conv3d_stem_wgrad(cl(x), cl(dy), B, D, H, W)
    assert_close(dw.cpu(), w.grad, 2e-5, "stem wgrad")


def test_conv3d_wgrad_two_sources_96():
    ops = _ops()
    B, cins, cout, D, H, W = 1, [16, 32], 16, 8, 24, 40
    xs = [torch.randn(B, c, D, H, W, generator=g(i)) for i, c in enumerate(cins)]
    dy = torch.randn(B, cout, D, H, W, generator=g(4))
    xcat = torch.cat(xs, 1).requires_grad_(True)
    w = torch.zeros(cout, 48, 3, 3, 3, requires_grad=True)
    F.conv3d(xcat, w, None, padding=1).backward(dy)
    dw, db = ops.conv3d_wgrad([cl(x) for x in xs], cins, cl(dy), cout, B, D, H, W)
    assert_close(dw.cpu(), w.grad, 1e-5, "wgrad 2src")
    assert_close(db.cpu(), dy.sum((0, 2, 3, 4)), 1e-5, "bgrad")


UMMA_CASES = [
    # B, cins, cout, (D,H,W)
    (1, [16], 16, (4, 16, 8)),      # exactly one tile per plane
    (2, [16], 16, (6, 32, 24)),     # several tiles, two samples
    (1, [32], 32, (5, 24, 24)),     # ragged H tile (24 = 16 + 8)
    (1, [64], 64, (3, 12, 12)),     # ragged H and W
    (1, [16, 32], 16, (4, 16, 16)), # virtual concat (up1 shape)
    (1, [128, 256], 128, (2, 12, 12)),
    (1, [128], 256, (2, 6, 6)),     # two N tiles
    (1, [256], 256, (3, 6, 6)),
]


@pytest.mark.parametrize("mode,tol", [("parity", 2e-4), ("fast", 3e-2)])
@pytest.mark.parametrize("B,cins,cout,dims", UMMA_CASES)
def test_conv3d_umma_fwd(B, cins, cout, dims, mode, tol):
    import icl_b200
    ops = _ops()
    icl_b200.set_precision(mode)
    try:
        D, H, W = dims
        xs = [torch.randn(B, c, D, H, W, generator=g(i)) for i, c in enumerate(cins)]
        w = torch.randn(cout, sum(cins), 3, 3, 3, generator=g(9)) * (2.0 / (27 * sum(cins))) ** 0.5
        b = torch.randn(cout, generator=g(10))
        ref = F.conv3d(torch.cat(xs, 1).double(), w.double(), b.double(), padding=1)
        pks = [ops.pack_pk(cl(x)) for x in xs]
        stats = torch.zeros(B, cout, 2, dtype=torch.float64, device="cuda")
        y = ops.conv3d_umma(pks, cins, ops.pack_w_umma(w.cuda(), False), b.cuda(), cout, B, D, H, W, stats)
        torch.cuda.synchronize()
        assert_close(uncl(y), ref, tol, "umma fwd %s" % mode)
        assert_close(stats[..., 0].cpu(), ref.sum((2, 3, 4)), tol, "sum", abs_floor=1e-2 if mode == "parity" else 1.0)
        assert_close(stats[..., 1].cpu(), (ref ** 2).sum((2, 3, 4)), tol * 3, "sumsq")
    finally:
        icl_b200.set_precision("parity")


WALK_CASES = [
    # B, cins, cout, (D,H,W)     (D >= 16 so that the plane-walk kernel is selected)
    (1, [16], 16, (16, 16, 8)),       # one column, one segment
    (2, [16], 16, (40, 32, 24)),      # ring of 32 slots wraps, several segments / samples
    (1, [16, 32], 16, (24, 16, 16)),  # virtual concat (up1 shape), 3 chunks
    (1, [16], 32, (18, 24, 12)),      # Cout 32: ring of 16 slots, ragged tiles
    (1, [32], 32, (20, 12, 12)),
]


@pytest.mark.parametrize("mode,tol", [("parity", 2e-4), ("fast", 3e-2)])
@pytest.mark.parametrize("B,cins,cout,dims", WALK_CASES)
def test_conv3d_umma_walk_fwd(B, cins, cout, dims, mode, tol):
    """plane-walk kernel (depth taps folded into N, TMEM accumulator ring) vs F.conv3d, incl. the fused statistics."""
    import icl_b200
    ops = _ops()
    icl_b200.set_precision(mode)
    try:
        D, H, W = dims
        xs = [torch.randn(B, c, D, H, W, generator=g(i)) for i, c in enumerate(cins)]
        w = torch.randn(cout, sum(cins), 3, 3, 3, generator=g(9)) * (2.0 / (27 * sum(cins))) ** 0.5
        b = torch.randn(cout, generator=g(10))
        ref = F.conv3d(torch.cat(xs, 1).double(), w.double(), b.double(), padding=1)
        pks = [ops.pack_pk(cl(x)) for x in xs]
        wp = ops.pack_w_umma(w.cuda(), False, D)
        assert isinstance(wp, tuple) and wp[0] == "walk"
        stats = torch.zeros(B, cout, 2, dtype=torch.float64, device="cuda")
        for max_ctas in (0, 3):  # 3 CTAs: every CTA walks many items, ring + stage phases wrap
            stats.zero_()
            old = ops._MAX_CTAS
            ops._MAX_CTAS = max_ctas
            try:
                y = ops.conv3d_umma(pks, cins, wp, b.cuda(), cout, B, D, H, W, stats)
            finally:
                ops._MAX_CTAS = old
            torch.cuda.synchronize()
            assert_close(uncl(y), ref, tol, "walk fwd %s" % mode)
            assert_close(stats[..., 0].cpu(), ref.sum((2, 3, 4)), tol, "sum", abs_floor=1e-2 if mode == "parity" else 1.0)
            assert_close(stats[..., 1].cpu(), (ref ** 2).sum((2, 3, 4)), tol * 3, "sumsq")
    finally:
        icl_b200.set_precision("parity")


def test_conv3d_umma_walk_dgrad_split():
    """data gradient of a concatenated input through the plane-walk kernel: N = 48 per plane, two outputs."""
    ops = _ops()
    B, cins, cout, D, H, W = 1, [16, 32], 16, 20, 16, 16
    dy = torch.randn(B, cout, D, H, W, generator=g(5))
    w = torch.randn(cout, 48, 3, 3, 3, generator=g(6)) * 0.05
    x = torch.zeros(B, 48, D, H, W, requires_grad=True)
    F.conv3d(x, w, None, padding=1).backward(dy)
    wp = ops.pack_w_umma(w.cuda(), True, D)
    assert isinstance(wp, tuple)
    d0, d1 = ops.conv3d_umma([ops.pack_pk(cl(dy))], [cout], wp, None, 48, B, D, H, W, split=16)
    assert_close(uncl(d0), x.grad[:, :16], 2e-4, "walk dgrad skip part")
    assert_close(uncl(d1), x.grad[:, 16:], 2e-4, "walk dgrad up part")


def test_conv3d_umma_persistent_loop_and_dgrad_split():
    """few CTAs => each walks many tiles (ring + accumulator phases wrap); dgrad writes two outputs."""
    ops = _ops()
    B, cins, cout, D, H, W = 2, [16, 32], 16, 6, 32, 16
    dy = torch.randn(B, cout, D, H, W, generator=g(5))
    w = torch.randn(cout, 48, 3, 3, 3, generator=g(6)) * 0.05
    x = torch.zeros(B, 48, D, H, W, requires_grad=True)
    F.conv3d(x, w, None, padding=1).backward(dy)
    old = ops._MAX_CTAS
    ops._MAX_CTAS = 3
    try:
        d0, d1 = ops.conv3d_umma([ops.pack_pk(cl(dy))], [cout], ops.pack_w_umma(w.cuda(), True), None, 48, B, D, H, W, split=16)
    finally:
        ops._MAX_CTAS = old
    assert_close(uncl(d0), x.grad[:, :16], 2e-4, "dgrad skip part")
    assert_close(uncl(d1), x.grad[:, 16:], 2e-4, "dgrad up part")


WGRAD_CASES = [
    # B, cins, cout, (D,H,W)
    (1, [16], 16, (2, 16, 8)),       # one window step, one tile
    (2, [16], 16, (6, 32, 24)),      # several steps / tiles / samples -> several splits
    (1, [32], 16, (4, 24, 12)),      # ragged tiles, two ci tiles
    (1, [16, 32], 16, (4, 16, 16)),  # virtual concat (up1 shape): two sources
    (1, [64], 64, (4, 12, 12)),
    (1, [128], 256, (2, 6, 6)),      # more output blocks than SMs
    (1, [16], 16, (3, 20, 10)),      # odd depth, ragged tiles in both directions (the tensor-memory kernel walks plane by plane)
    (2, [32], 32, (1, 16, 8)),       # a single plane: both depth neighbours are out of range
    (1, [48], 16, (5, 8, 8)),        # Cin not a multiple of 32: three 16-channel blocks
]


@pytest.mark.parametrize("mode,tol", [("parity", 2e-4), ("fast", 3e-2)])
@pytest.mark.parametrize("B,cins,cout,dims", WGRAD_CASES)
def test_conv3d_wgrad_umma(B, cins, cout, dims, mode, tol):
    import icl_b200
    ops = _ops()
    icl_b200.set_precision(mode)
    try:
        D, H, W = dims
        xs = [torch.randn(B, c, D, H, W, generator=g(i)) for i, c in enumerate(cins)]
        dy = torch.randn(B, cout, D, H, W, generator=g(4))
        xcat = torch.cat(xs, 1).double().requires_grad_(True)
        w = torch.zeros(cout, sum(cins), 3, 3, 3, dtype=torch.float64, requires_grad=True)
        F.conv3d(xcat, w, None, padding=1).backward(dy.double())
        if D % 2 and not ops.wgrad_ts():
            pytest.skip("the shared-memory-operand kernel (ICL_WGRAD=smem) needs an even depth")
        dw = ops.conv3d_wgrad_umma([ops.pack_pk(cl(x)) for x in xs], cins, ops.pack_pk(cl(dy)), cout, B, D, H, W)
        torch.cuda.synchronize()
        assert_close(dw.cpu(), w.grad, tol, "wgrad umma %s" % mode)
        dw2 = ops.conv3d_wgrad_umma([ops.pack_pk(cl(x)) for x in xs], cins, ops.pack_pk(cl(dy)), cout, B, D, H, W)
        torch.cuda.synchronize()
        assert torch.equal(dw, dw2), "the weight gradient must be bit-reproducible (fixed-order reduction of the per-CTA partial sums)"
    finally:
        icl_b200.set_precision("parity")


# ------------------------------------------------------------------------------------------ norm / pool / upsample / dropout
@pytest.mark.parametrize("C", [16, 5])
def test_instnorm_relu_fwd_bwd(C):
    ops = _ops()
    B, D, H, W = 2, 6, 7, 9
    y = (torch.randn(B, C, D, H, W, generator=g(1)) * 2 + 0.5).requires_grad_(True)
    dA = torch.randn(B, C, D, H, W, generator=g(2))
    a_ref = F.relu(F.instance_norm(y, eps=1e-5))
    a_ref.backward(dA)
    yc = cl(y.detach())
    stats = torch.zeros(B, C, 2, dtype=torch.float64, device="cuda")
    ops.call("icl_instnorm_stats", ops.P(yc), ops.P(stats), ops.c_int(B), ops.c_int(C), ops.c_ll(D * H * W))
    mr = ops.instnorm_finalize(stats, B, C, D * H * W)
    a, pk = ops.instnorm_relu_fwd(yc, mr, C % 16 == 0)
    assert_close(uncl(a), a_ref.detach(), 1e-5, "IN+ReLU fwd")
    if pk is not None:
        rec = (pk[0].float() + pk[1].float()).permute(0, 1, 5, 2, 3, 4).reshape(B, C, D, H, W).cpu()
        assert_close(rec, a_ref.detach(), 2e-5, "PK hi+lo")
    if C % 8 == 0:
        dY, dpk, db = ops.instnorm_relu_bwd(cl(dA), yc, mr, C % 16 == 0, want_dbias=True)
        # the bias gradient in front of InstanceNorm is mathematically zero: compare with an absolute floor
        assert_close(db.cpu(), y.grad.sum((0, 2, 3, 4)), 1e-3, "dbias", abs_floor=1e-4)
    else:
        dY, dpk = ops.instnorm_relu_bwd(cl(dA), yc, mr, False)
    assert_close(uncl(dY), y.grad, 2e-5, "IN+ReLU bwd")
    if dpk is not None:
        rec = (dpk[0].float() + dpk[1].float()).permute(0, 1, 5, 2, 3, 4).reshape(B, C, D, H, W).cpu()
        assert_close(rec, y.grad, 3e-5, "PK dY")


def test_maxpool_ties_first_max():
    ops = _ops()
    B, C, D, H, W = 2, 16, 4, 6, 8
    x = torch.randint(0, 3, (B, C, D, H, W), generator=g(3)).float()  # many exact ties, like post-ReLU zeros
    xr = x.clone().requires_grad_(True)
    out_ref = F.max_pool3d(xr, 2)
    dout = torch.randn(out_ref.shape, generator=g(4))
    out_ref.backward(dout)
    out, idx, pk = ops.maxpool_fwd(cl(x), True)
    assert torch.equal(uncl(out), out_ref.detach())
    rec = (pk[0].float() + pk[1].float()).permute(0, 1, 5, 2, 3, 4).reshape(B, C, D // 2, H // 2, W // 2).cpu()
    assert torch.equal(rec, out_ref.detach()), "maxpool PK hi+lo"
    dx = torch.empty(B, D, H, W, C, device="cuda")
    ops.maxpool_bwd(cl(dout), idx, dx, False)
    assert torch.equal(uncl(dx), xr.grad)
    base = torch.randn(B, D, H, W, C, generator=g(5)).cuda()
    dx2 = base.clone()
    ops.maxpool_bwd(cl(dout), idx, dx2, True)
    assert_close(dx2.cpu(), (base + cl(xr.grad)).cpu(), 1e-6, "accumulate")


@pytest.mark.parametrize("C", [16, 6, 64, 24])
def test_upsample2x_fwd_bwd(C):
    ops = _ops()
    B, d, h, w = 2, 3, 4, 5
    x = torch.randn(B, C, d, h, w, generator=g(1), requires_grad=True)
    up = F.interpolate(x, scale_factor=(2, 2, 2), mode="trilinear", align_corners=False)
    dout = torch.randn(up.shape, generator=g(2))
    up.backward(dout)
    out, pk = ops.upsample2x_fwd(cl(x.detach()), C % 8 == 0)
    assert_close(uncl(out), up.detach(), 1e-6, "upsample fwd")
    if pk is not None:
        none, pk2 = ops.upsample2x_fwd(cl(x.detach()), True, want_f32=False)  # PK-only variant used by the lean backbone
        assert none is None and torch.equal(pk, pk2)
        rec = (pk[0].float() + pk[1].float()).permute(0, 1, 5, 2, 3, 4).reshape(B, C, 2 * d, 2 * h, 2 * w).cpu()
        assert_close(rec, up.detach(), 2e-5, "upsample PK hi+lo")
    dx = torch.empty(B, d, h, w, C, device="cuda")
    ops.upsample2x_bwd(cl(dout), 0, C, dx, False)
    assert_close(uncl(dx), x.grad, 1e-5, "upsample bwd")
    # channel-offset form used for the [skip|up] gradient
    wide = torch.cat([torch.randn(B, 8, 2 * d, 2 * h, 2 * w, generator=g(3)), dout], 1)
    dx2 = torch.empty(B, d, h, w, C, device="cuda")
    ops.upsample2x_bwd(cl(wide), 8, C, dx2, False)
    assert_close(uncl(dx2), x.grad, 1e-5, "upsample bwd offset")


def test_dropout_mask_and_philox():
    ops = _ops()
    x = torch.randn(2, 6, 6, 6, 16, generator=g(1)).cuda()
    mask = (torch.rand(x.shape, generator=g(2)) > 0.3).to(torch.uint8).cuda()
    out = ops.dropout(x, 0.3, mask, 0)
    assert_close(out.cpu(), (x * mask.float() / 0.7).cpu(), 1e-6, "dropout mask")
    big = torch.ones(1 << 20, device="cuda")
    o1 = ops.dropout(big, 0.3, None, 1234)
    o2 = ops.dropout(big, 0.3, None, 1234)
    o3 = ops.dropout(big, 0.3, None, 1235)
    assert torch.equal(o1, o2) and not torch.equal(o1, o3)
    keep = (o1 != 0).float().mean().item()
    assert abs(keep - 0.7) < 5e-3, keep
    assert abs(o1.max().item() - 1 / 0.7) < 1e-6


TOK_CASES = [(3136, 288, 96, 1), (1000, 64, 64, 0), (2048, 96, 384, 0), (1024, 20, 36, 1), (4096, 384, 96, 1)]


@pytest.mark.parametrize("M,N,K,act", TOK_CASES)
def test_tok_linear_umma(monkeypatch, M, N, K, act):
    """Token-major nn.Linear on the tcgen05 kernel (Swin qkv / proj / mlp, ICL token projections): forward with bias / GELU and the
    pre-activation copy, the data gradient and the weight gradient, against float64 (split-bf16 products: 1e-4)."""
    monkeypatch.setenv("ICL_TOKLIN_MIN_FLOP", "0")   # small test shapes: lift the size threshold of the routing
    ops = _ops()
    assert ops.tok_linear_ok(M, N, K)
    x = torch.randn(M, K, generator=g(11)).cuda()
    w = (torch.randn(N, K, generator=g(12)) * 0.2).cuda()
    b = torch.randn(N, generator=g(13)).cuda()
    y, pre = ops.linear_fwd(x, w, b, act, want_pre=bool(act))
    ref_pre = x.double().cpu() @ w.double().cpu().t() + b.double().cpu()
    ref = F.gelu(ref_pre) if act else ref_pre
    assert_close(y.cpu(), ref, 1e-4, "tok linear fwd")
    if act:
        assert_close(pre.cpu(), ref_pre, 1e-4, "tok linear pre-activation")
    dy = torch.randn(M, N, generator=g(14)).cuda()
    if ops.tok_linear_ok(M, K, N):
        dx = ops.linear_dgrad(dy, w)
        assert_close(dx.cpu(), dy.double().cpu() @ w.double().cpu(), 1e-4, "tok linear dgrad")
    if M >= 2048:
        dW, db = ops.linear_wgrad(dy, x, True)
        assert_close(dW.cpu(), dy.double().cpu().t() @ x.double().cpu(), 1e-4, "tok linear wgrad")
        assert_close(db.cpu(), dy.double().cpu().sum(0), 1e-4, "tok linear bias grad")


# ------------------------------------------------------------------------------------------ GEMM family / heads
@pytest.mark.parametrize("M,N,K", [(7, 33, 19), (130, 70, 65), (16, 2048, 1024), (3, 1030, 1100), (40, 1536, 1536), (32, 1728, 1728),
                                   (16, 1100, 2052), (64, 1027, 1028), (16, 4100, 2052), (24, 4608, 1028), (5, 4097, 260), (16, 8200, 2052),
                                   (24, 4608, 4100), (40, 4224, 4100), (100, 4224, 4100), (128, 4100, 4224)])
def test_linear_fwd_bwd(M, N, K):
    import icl_b200.functional as Fn
    x = torch.randn(M, K, generator=g(1), requires_grad=True)
    w = (torch.randn(N, K, generator=g(2)) / K ** 0.5).requires_grad_(True)
    b = torch.randn(N, generator=g(3)).requires_grad_(True)
    for act in (0, 1):
        ref = F.linear(x, w, b)
        ref = F.gelu(ref) if act else ref
        dy = torch.randn(M, N, generator=g(4))
        gx, gw, gb = torch.autograd.grad(ref, (x, w, b), dy)
        xc, wc, bc = (t.detach().cuda().requires_grad_(True) for t in (x, w, b))
        y = Fn.linear(xc, wc, bc, act)
        y.backward(dy.cuda())
        assert_close(y.detach().cpu(), ref.detach(), 2e-5, "linear fwd act=%d" % act)
        assert_close(xc.grad.cpu(), gx, 2e-5, "linear dx")
        assert_close(wc.grad.cpu(), gw, 2e-5, "linear dw")
        assert_close(bc.grad.cpu(), gb, 2e-5, "linear db")


def test_skinny_linear_abi_row_blocks():
    """The C-ABI skinny GEMMs take up to 256 rows (one weight pass per 64-row block); Python routes > 64 rows elsewhere."""
    from icl_b200.ops import P, c_int, call
    M, N, K = 100, 1100, 1028
    x = torch.randn(M, K, generator=g(1)).cuda()
    w = (torch.randn(N, K, generator=g(2)) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g(3)).cuda()
    y = torch.empty(M, N, device="cuda")
    call("icl_skinny_linear_fwd", c_int(M), c_int(N), c_int(K), P(x), P(w), P(b), P(y), P(None), c_int(0))
    assert_close(y.cpu(), F.linear(x.double(), w.double(), b.double()).cpu(), 2e-5, "skinny fwd 100 rows")
    dy = torch.randn(M, N, generator=g(4)).cuda()
    dx = torch.zeros(M, K, device="cuda")
    call("icl_skinny_linear_dgrad", c_int(M), c_int(N), c_int(K), P(dy), P(w), P(dx))
    assert_close(dx.cpu(), (dy.double() @ w.double()).cpu(), 2e-5, "skinny dgrad 100 rows")


@pytest.mark.parametrize("M,N,K", [(100000, 16, 2), (70001, 64, 64), (33000, 128, 64)])
def test_linear_long_reduction_splitk(M, N, K):
    """1x1x1-conv / Linear weight and bias gradients reduce over all voxels: split-K sgemm + parallel column sums
    (the reference shapes are final: rows = B*96^3, proj/fc_kv: rows = B*24^3)."""
    import icl_b200.functional as Fn
    x = torch.randn(M, K, generator=g(1), requires_grad=True)
    w = (torch.randn(N, K, generator=g(2)) / K ** 0.5).requires_grad_(True)
    b = torch.randn(N, generator=g(3)).requires_grad_(True)
    dy = torch.randn(M, N, generator=g(4))
    gx, gw, gb = torch.autograd.grad(F.linear(x.double(), w.double(), b.double()), (x, w, b), dy.double())
    xc, wc, bc = (t.detach().cuda().requires_grad_(True) for t in (x, w, b))
    Fn.linear(xc, wc, bc, 0).backward(dy.cuda())
    assert_close(xc.grad.cpu(), gx, 2e-5, "dx")
    assert_close(wc.grad.cpu(), gw, 2e-5, "dw split-K")
    assert_close(bc.grad.cpu(), gb, 2e-5, "db split rows")


@pytest.mark.parametrize("rows,C", [(10, 64), (6, 1728), (33, 100)])
def test_layernorm(rows, C):
    import icl_b200.functional as Fn
    x = (torch.randn(rows, C, generator=g(1)) * 3 + 1).requires_grad_(True)
    w = (1 + 0.1 * torch.randn(C, generator=g(2))).requires_grad_(True)
    b = (0.1 * torch.randn(C, generator=g(3))).requires_grad_(True)
    dy = torch.randn(rows, C, generator=g(4))
    ref = F.layer_norm(x, (C,), w, b, 1e-5)
    gx, gw, gb = torch.autograd.grad(ref, (x, w, b), dy)
    xc, wc, bc = (t.detach().cuda().requires_grad_(True) for t in (x, w, b))
    y = Fn.layer_norm(xc, wc, bc)
    y.backward(dy.cuda())
    assert_close(y.detach().cpu(), ref.detach(), 1e-5, "ln fwd")
    assert_close(xc.grad.cpu(), gx, 3e-5, "ln dx")
    assert_close(wc.grad.cpu(), gw, 3e-5, "ln dw")
    assert_close(bc.grad.cpu(), gb, 3e-5, "ln db")


# head_dim 16 takes the split-N kernels (N = 2500: three voxel chunks); (1, 300, 64, 2, 3) has head_dim 32: the generic kernels
@pytest.mark.parametrize("B,N,C,H,K", [(2, 216, 64, 4, 3), (1, 500, 32, 2, 16), (2, 64, 16, 1, 2), (2, 2500, 64, 4, 2), (1, 3000, 32, 2, 16),
                                       (1, 300, 64, 2, 3)])
def test_proxy_attention(B, N, C, H, K):
    import icl_b200.functional as Fn
    hd = C // H
    ql = torch.randn(B, K, C, generator=g(1), requires_grad=True)
    kv = torch.randn(B, N, 2 * C, generator=g(2), requires_grad=True)
    qh = ql.reshape(B, H, K, hd)
    k = kv[..., :C].reshape(B, N, H, hd)
    v = kv[..., C:].reshape(B, N, H, hd)
    logits = torch.einsum("bhkd,bnhd->bhkn", qh, k) * hd ** -0.5
    xv_ref = torch.einsum("bhkn,bnhd->bhkd", logits.softmax(-1), v).reshape(B, K, C)
    map_ref = logits.permute(0, 2, 1, 3)
    dxv = torch.randn(B, K, C, generator=g(3))
    dmap = torch.randn(B, K, H, N, generator=g(4))
    gq, gkv = torch.autograd.grad((xv_ref, map_ref), (ql, kv), (dxv, dmap), retain_graph=True)
    qc, kc = ql.detach().cuda().requires_grad_(True), kv.detach().cuda().requires_grad_(True)
    xv, amap = Fn.proxy_attention(qc, kc, H)
    torch.autograd.backward((xv, amap), (dxv.cuda(), dmap.cuda()))
    assert_close(xv.detach().cpu(), xv_ref.detach(), 2e-5, "xv")
    assert_close(amap.detach().cpu(), map_ref.detach(), 1e-5, "map")
    assert_close(qc.grad.cpu(), gq, 5e-5, "dq")
    assert_close(kc.grad.cpu(), gkv, 5e-5, "dkv")
    # map-only gradient (the uscl branch): v-half of dkv must be exactly zero
    qc2, kc2 = ql.detach().cuda().requires_grad_(True), kv.detach().cuda().requires_grad_(True)
    _, amap2 = Fn.proxy_attention(qc2, kc2, H, want_xv=False)
    amap2.backward(dmap.cuda())
    gq2, gkv2 = torch.autograd.grad(map_ref, (ql, kv), dmap)
    assert_close(qc2.grad.cpu(), gq2, 5e-5, "dq map-only")
    assert_close(kc2.grad.cpu(), gkv2, 5e-5, "dkv map-only")
    assert kc2.grad[..., C:].abs().max().item() == 0.0


def test_separable_conv_stack():
    import icl_b200.functional as Fn
    NB, CH, d = 6, 4, 5
    x = torch.randn(NB, CH, d, d, d, generator=g(1), requires_grad=True)
    wd = (torch.randn(CH, 1, 3, 3, 3, generator=g(2)) * 0.3).requires_grad_(True)
    wp = (torch.randn(CH, CH, 1, 1, 1, generator=g(3)) * 0.5).requires_grad_(True)
    w1 = (torch.randn(1, CH, 1, 1, 1, generator=g(4)) * 0.5).requires_grad_(True)
    b1 = torch.randn(1, generator=g(5)).requires_grad_(True)
    bn1, bn2 = torch.nn.BatchNorm3d(CH), torch.nn.BatchNorm3d(CH)
    for bn, s in ((bn1, 6), (bn2, 7)):
        with torch.no_grad():
            bn.weight.copy_(1 + 0.2 * torch.randn(CH, generator=g(s)))
            bn.bias.copy_(0.2 * torch.randn(CH, generator=g(s + 10)))
    import copy
    cbn1, cbn2 = copy.deepcopy(bn1).cuda(), copy.deepcopy(bn2).cuda()
    y = F.relu(bn1(F.conv3d(x, wd, None, 1, 1, 1, CH)))
    y = F.relu(bn2(F.conv3d(y, wp)))
    y = F.conv3d(y, w1, b1)
    dy = torch.randn(y.shape, generator=g(8))
    params = (x, wd, wp, w1, b1, bn1.weight, bn1.bias, bn2.weight, bn2.bias)
    grads = torch.autograd.grad(y, params, dy)
    xc, wdc, wpc, w1c, b1c = (t.detach().cuda().requires_grad_(True) for t in (x, wd, wp, w1, b1))
    z = Fn.bn_relu(Fn.dwconv3d(xc, wdc), cbn1, True)
    z = Fn.bn_relu(Fn.planar_pointwise(z, wpc, None), cbn2, True)
    z = Fn.planar_pointwise(z, w1c, b1c)
    z.backward(dy.cuda())
    assert_close(z.detach().cpu(), y.detach(), 2e-5, "sepconv fwd")
    mine = (xc.grad, wdc.grad, wpc.grad, w1c.grad, b1c.grad, cbn1.weight.grad, cbn1.bias.grad, cbn2.weight.grad, cbn2.bias.grad)
    for name, a, b in zip("x wd wp w1 b1 g1 be1 g2 be2".split(), mine, grads):
        assert_close(a.cpu(), b, 1e-4, "sepconv d" + name, abs_floor=1e-6)
    assert_close(cbn1.running_mean.cpu(), bn1.running_mean, 1e-5, "running_mean")
    assert_close(cbn2.running_var.cpu(), bn2.running_var, 1e-5, "running_var")
    assert int(cbn1.num_batches_tracked) == 1


def test_add_scaled_and_batch_mean():
    import icl_b200.functional as Fn
    a = torch.randn(3, 4, 5, generator=g(1), requires_grad=True)
    r = torch.tensor([0.0, 1 / 0.98, 1 / 0.98])
    ref = a + a * r.view(3, 1, 1)
    m_ref = ref.mean(0, keepdim=True)
    dm = torch.randn(1, 4, 5, generator=g(2))
    (ga,) = torch.autograd.grad(m_ref, a, dm)
    ac = a.detach().cuda().requires_grad_(True)
    out = Fn.batch_mean(Fn.add_scaled(ac, ac, r.cuda()))
    out.backward(dm.cuda())
    assert_close(out.detach().cpu(), m_ref.detach(), 1e-6, "mean")
    assert_close(ac.grad.cpu(), ga, 1e-6, "grad")


# ------------------------------------------------------------------------------------------ losses / optimiser / inference kernels
@pytest.mark.parametrize("K,generic", [(2, False), (5, False), (16, False), (4, True), (16, True)])
def test_losses_small(monkeypatch, K, generic):
    """generic=True routes the class-statistics losses to the per-voxel kernels (grids whose rows exceed the row kernels' staging);
    the default is the row form (z/y interpolation table in shared memory, x-adjoint inside the block)."""
    from icl_b200.utils import losses as L
    from oracle import restate as R
    monkeypatch.setenv("ICL_DISABLE_ROW_LOSS", "1" if generic else "0")
    B, S = 2, 16
    size = (S, S, S)
    labels = torch.randint(0, K, (B, S, S, S), generator=g(1))
    logits = torch.randn(B, K, S, S, S, generator=g(2), requires_grad=True)
    unl = torch.randn(B, K, S, S, S, generator=g(3))
    fms = [(torch.randn(B, K, r, r, r, generator=g(10 + r)) * 2).requires_grad_(True) for r in (2, 4, 8)]
    fms2 = [(torch.randn(B, K, r, r, r, generator=g(20 + r)) * 2).requires_grad_(True) for r in (2, 4, 8)]
    fms3 = [torch.randn(B, K, r, r, r, generator=g(30 + r)) * 2 for r in (2, 4, 8)]
    ref = R.icl_losses((logits, unl, fms, fms2, fms3), labels, K, size=size)
    ref["total"].backward()
    lc = logits.detach().cuda().permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3).requires_grad_(True)
    f1 = [t.detach().cuda().requires_grad_(True) for t in fms]
    f2 = [t.detach().cuda().requires_grad_(True) for t in fms2]
    f3 = [t.cuda() for t in fms3]
    yl = labels.cuda()
    ce = L.CrossEntropyLoss()(lc, yl)
    dice = L.DiceLoss(K)(torch.softmax(lc, 1), yl.unsqueeze(1))
    aux = L.AuxLoss3D(K, size)(f1, yl)
    pse = L.PseudoSoftLoss3D(K, size)(f2, unl.cuda())
    cons = L.softmax_mse_loss(f2, f3)
    total = dice + ce + aux + pse + 10 * cons
    total.backward()
    for name, v in (("ce", ce), ("dice", dice), ("aux", aux), ("pse", pse), ("cons", cons), ("total", total)):
        assert abs(v.item() - ref[name].item()) <= 2e-5 * max(1.0, abs(ref[name].item())), (name, v.item(), ref[name].item())
    assert_close(lc.grad.cpu(), logits.grad, 1e-4, "dlogits")
    for i in range(3):
        assert_close(f1[i].grad.cpu(), fms[i].grad, 1e-4, "daux%d" % i)
        assert_close(f2[i].grad.cpu(), fms2[i].grad, 1e-4, "dpse%d" % i)
    # DiceLoss(softmax=True) + class weights, fused CE+Dice on logits
    w = [0.5 + 0.1 * k for k in range(K)]
    l2 = logits.detach().clone().requires_grad_(True)
    soft = torch.softmax(l2, 1)
    refw = sum(w[k] * (1 - (2 * (soft[:, k] * (labels == k)).sum() + 1e-5) / ((soft[:, k] ** 2).sum() + (labels == k).float().sum() + 1e-5))
               for k in range(K)) / K
    refw.backward()
    l3 = logits.detach().cuda().requires_grad_(True)
    mine = L.DiceLoss(K)(l3, yl.unsqueeze(1), weight=w, softmax=True)
    mine.backward()
    assert abs(mine.item() - refw.item()) < 2e-5
    assert_close(l3.grad.cpu(), l2.grad, 1e-4, "weighted dice grad")


@pytest.mark.parametrize("K", [2, 4, 16])
def test_head1x1_fwd_bwd(K):
    """`final` 1x1x1 conv (16 -> K) streaming kernels vs F.conv3d and its autograd."""
    ops = _ops()
    rows = 70003
    x = torch.randn(rows, 16, generator=g(1), requires_grad=True)
    w = (torch.randn(K, 16, generator=g(2)) * 0.3).requires_grad_(True)
    b = torch.randn(K, generator=g(3)).requires_grad_(True)
    gy = torch.randn(rows, K, generator=g(4))
    ref = F.linear(x, w, b)
    ref.backward(gy)
    out = torch.empty(rows, K, device="cuda")
    xc, wc, bc, gc = x.detach().cuda(), w.detach().cuda(), b.detach().cuda(), gy.cuda()  # keep the device buffers alive
    ops.call("icl_head1x1_fwd", ops.P(xc), ops.P(wc), ops.P(bc), ops.P(out), ops.c_ll(rows), ops.c_int(16), ops.c_int(K))
    assert_close(out.cpu(), ref.detach(), 1e-6, "head fwd")
    if K in (2, 4, 16):
        dx = torch.empty(rows, 16, device="cuda")
        dw, db = torch.zeros(K, 16, device="cuda"), torch.zeros(K, device="cuda")
        ops.call("icl_head1x1_bwd", ops.P(gc), ops.P(xc), ops.P(wc), ops.P(dx), ops.P(dw), ops.P(db), ops.c_ll(rows), ops.c_int(16), ops.c_int(K))
        assert_close(dx.cpu(), x.grad, 1e-6, "head dx")
        assert_close(dw.cpu(), w.grad, 2e-5, "head dw")
        assert_close(db.cpu(), b.grad, 2e-5, "head db")


def test_sgd_multi_matches_torch():
    from icl_b200.optim import SGD
    shapes = [(5,), (33, 7), (70000,), (16, 16, 3, 3, 3), (1,)]
    ps = [torch.randn(s, generator=g(i)) for i, s in enumerate(shapes)]
    ref = [p.clone().requires_grad_(True) for p in ps]
    mine = [p.clone().cuda().requires_grad_(True) for p in ps]
    o_ref = torch.optim.SGD(ref, lr=0.01, momentum=0.9, weight_decay=1e-4)
    o_mine = SGD(mine, lr=0.01, momentum=0.9, weight_decay=1e-4)
    for step in range(3):
        for i, (a, b) in enumerate(zip(ref, mine)):
            if i == 4 and step < 2:  # a parameter whose grad stays None is skipped entirely
                a.grad = None
                b.grad = None
                continue
            gr = torch.randn(a.shape, generator=g(100 + 10 * step + i))
            a.grad = gr.clone()
            b.grad = gr.cuda()
        o_ref.step()
        o_mine.step()
        lr = 0.01 * (1 - (step + 1) / 10) ** 0.9
        for grp in o_ref.param_groups:
            grp["lr"] = lr
        for grp in o_mine.param_groups:
            grp["lr"] = lr
    for a, b in zip(ref, mine):
        assert_close(b.detach().cpu(), a.detach(), 1e-6, "sgd param")


def test_sgd_fused_factored_matches_torch():
    """fused rank-R weight gradient + SGD (mlp2 path): a Linear whose weight is above the factoring threshold is used twice per
    step (like mlp2 in the sspa and uscl passes); .grad is never materialised, the update equals torch.optim.SGD's."""
    import icl_b200.functional as Fn
    from icl_b200.optim import SGD
    N, K, M = 70, 52, 9
    w0 = torch.randn(N, K, generator=g(1)) * 0.2
    b0 = torch.randn(N, generator=g(2))
    wr, br = w0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
    wm, bm = w0.clone().cuda().requires_grad_(True), b0.clone().cuda().requires_grad_(True)
    o_ref = torch.optim.SGD([wr, br], lr=0.05, momentum=0.9, weight_decay=1e-4)
    o_mine = SGD([wm, bm], lr=0.05, momentum=0.9, weight_decay=1e-4, fused_factored=True, factored_min_numel=1000)
    for step in range(3):
        xa, xb = torch.randn(M, K, generator=g(10 + step)), torch.randn(M + 3, K, generator=g(20 + step))
        o_ref.zero_grad()
        (F.linear(xa, wr, br).pow(2).mean() + F.gelu(F.linear(xb, wr, br)).sum() * 0.01).backward()
        o_ref.step()
        o_mine.zero_grad()
        (Fn.linear(xa.cuda(), wm, bm).pow(2).mean() + Fn.linear(xb.cuda(), wm, bm, act=1).sum() * 0.01).backward()
        assert wm.grad is None and bm.grad is not None
        o_mine.step()
    assert_close(wm.detach().cpu(), wr.detach(), 2e-6, "fused factored weight")
    assert_close(bm.detach().cpu(), br.detach(), 2e-6, "bias")


def test_sliding_window_kernels_and_dice_counts():
    from icl_b200 import inference
    K, W, H, D = 3, 20, 18, 17
    score = torch.zeros(K, W, H, D, device="cuda")
    cnt = torch.zeros(W, H, D, device="cuda")
    ref_s = np.zeros((K, W, H, D), np.float32)
    ref_c = np.zeros((W, H, D), np.float32)
    for n, (xs, ys, zs) in enumerate([(0, 0, 0), (4, 2, 1), (4, 2, 1)]):
        lg = torch.randn(16, 16, 16, K, generator=g(n))
        inference._accumulate(lg.cuda(), score, cnt, xs, ys, zs)
        p = torch.softmax(lg, -1).permute(3, 0, 1, 2).numpy()
        ref_s[:, xs:xs + 16, ys:ys + 16, zs:zs + 16] += p
        ref_c[xs:xs + 16, ys:ys + 16, zs:zs + 16] += 1
    assert_close(score.cpu(), ref_s, 1e-6, "score")
    assert np.array_equal(cnt.cpu().numpy(), ref_c)
    cnt.clamp_(min=1)  # unvisited voxels exist only in this synthetic test
    lab = inference._finalize(score, cnt)
    ref_lab = np.argmax(ref_s / np.maximum(ref_c, 1)[None], axis=0)
    assert (lab.cpu().numpy() == ref_lab).mean() > 0.9999
    pred = torch.randint(0, 2, (W, H, D), generator=g(7))
    gt = torch.randint(0, 3, (W, H, D), generator=g(8))
    dice, counts = inference.dice_metric(pred.cuda(), gt.cuda())
    from oracle import restate as R
    rd, rc = R.dice_metric(pred.numpy(), gt.numpy())
    assert counts == rc and dice == rd
    assert inference.dice_metric(torch.zeros(4, 4, 4, dtype=torch.long).cuda(), torch.zeros(4, 4, 4, dtype=torch.long).cuda())[0] == 1.0


# ------------------------------------------------------------------------------------------ tcgen05 weight-streaming GEMMs
@pytest.mark.parametrize("M,N,K", [(16, 256, 128), (16, 4100, 2052), (48, 1000, 4224), (128, 4224, 4100), (100, 13824, 2048), (200, 1100, 1028),
                                   (300, 640, 512)])
def test_bigw_linear_abi(M, N, K):
    """icl_bigw_linear_{fwd,dgrad} straight through the C-ABI (tcgen05, W converted fp32 -> split bf16 into TMEM): edge tiles in the
    output axis, a reduction tail (K % 32 != 0), row padding (M % 16 != 0) and more than one 128-row pass."""
    from icl_b200 import _lib
    from icl_b200.ops import P, c_int, call
    x = torch.randn(M, K, generator=g(1)).cuda()
    w = (torch.randn(N, K, generator=g(2)) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g(3)).cuda()
    lib = _lib.lib()
    for act in (0, 1):
        y = torch.empty(M, N, device="cuda")
        pre = torch.empty(M, N, device="cuda")
        ws = torch.empty(int(lib.icl_bigw_workspace(M, N, K)), dtype=torch.uint8, device="cuda")
        call("icl_bigw_linear_fwd", c_int(M), c_int(N), c_int(K), P(x), P(w), P(b), P(y), P(pre), c_int(act), P(ws))
        ref = F.linear(x.double(), w.double(), b.double())
        assert_close(pre.cpu(), ref.cpu(), 2e-5, "bigw fwd pre-activation")
        assert_close(y.cpu(), (F.gelu(ref) if act else ref).cpu(), 2e-5, "bigw fwd act=%d" % act)
    dy = torch.randn(M, N, generator=g(4)).cuda()
    dx = torch.empty(M, K, device="cuda")
    ws = torch.empty(int(lib.icl_bigw_workspace(M, K, N)), dtype=torch.uint8, device="cuda")
    call("icl_bigw_linear_dgrad", c_int(M), c_int(N), c_int(K), P(dy), P(w), P(dx), P(ws))
    assert_close(dx.cpu(), (dy.double() @ w.double()).cpu(), 2e-5, "bigw dgrad")


@pytest.mark.parametrize("R1,R2,N,K,ctas", [(16, 0, 256, 512, 0), (9, 12, 1000, 776, 3), (128, 128, 2048, 1024, 0), (40, 300, 392, 264, 0)])
def test_sgd_factored_umma_abi(R1, R2, N, K, ctas):
    """icl_sgd_factored_pack / _apply through the C-ABI against torch.optim.SGD on the materialised gradient: two factor pairs (one
    scaled), ragged R, edge tiles in both weight axes, the persistent tile loop (ctas = 3), three steps (momentum)."""
    from icl_b200 import _lib
    from icl_b200.ops import P, c_f, c_int, call
    lib = _lib.lib()
    p0 = torch.randn(N, K, generator=g(1)) * 0.3
    ref = p0.clone().double().requires_grad_(True)
    o_ref = torch.optim.SGD([ref], lr=0.05, momentum=0.9, weight_decay=1e-2)
    p = p0.clone().cuda()
    m = torch.zeros_like(p)
    lr = torch.tensor([0.05], device="cuda")
    for step in range(3):
        pairs = [(torch.randn(R1, N, generator=g(10 + step)), torch.randn(R1, K, generator=g(20 + step)), 1.0)]
        if R2:
            pairs.append((torch.randn(R2, N, generator=g(30 + step)), torch.randn(R2, K, generator=g(40 + step)), 0.5))
        ref.grad = sum(s * (dy.double().t() @ x.double()) for dy, x, s in pairs)
        o_ref.step()
        R = R1 + R2
        ws = torch.empty(int(lib.icl_sgd_factored_workspace(R, N, K)), dtype=torch.uint8, device="cuda")
        r0 = 0
        for dy, x, s in pairs:
            dyc, xc = dy.cuda(), x.cuda()
            call("icl_sgd_factored_pack", P(dyc), P(xc), c_int(dy.shape[0]), c_int(r0), c_int(R), c_int(N), c_int(K), c_f(s), P(ws))
            r0 += dy.shape[0]
        call("icl_sgd_factored_apply", c_int(R), c_int(N), c_int(K), P(ws), P(p), P(m), P(lr), c_f(0.9), c_f(1e-2), c_int(ctas))
    assert_close(p.cpu(), ref.detach().float(), 2e-5, "sgd_factored_umma param")
