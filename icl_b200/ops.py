"""Thin Python wrappers over the C-ABI kernels (include/icl_b200.h).

Internal activation formats:
  NDHWC : contiguous float tensor [B, D, H, W, C]            (F32CL in the header)
  PK    : contiguous bf16 tensor  [P, B, C/8, D, H, W, 8]    (split hi/lo tensor-core operand)
At the module boundary NDHWC tensors are exposed as [B, C, D, H, W] views (channels_last_3d).
"""
import ctypes
import os

import torch

from . import _lib
from .precision import planes, tensor_cores

c_int, c_ll, c_f, c_d, c_ull = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_double, ctypes.c_ulonglong
P = _lib.ptr
call = _lib.call
profile_start, profile_stop = _lib.profile_start, _lib.profile_stop


def _require_cuda(t):
    if not t.is_cuda:
        raise RuntimeError("icl_b200 ops run on CUDA tensors only (no CPU fallback); got a %s tensor" % t.device)


def reduce_ws(device):
    """Scratch for the chunked two-stage reductions of the ICL-head kernels (heads.cu); allocated per call from torch's caching
    allocator (stream-ordered, CUDA-graph safe)."""
    return torch.empty((_lib.lib().icl_reduce_workspace_bytes(),), dtype=torch.uint8, device=device)


# Small zero-initialised buffers (statistics, atomics targets, bias gradients: ~200 per training step) are carved from a shared
# zeroed chunk instead of one fill kernel each.  A region is handed out once and never again; a full chunk is replaced by a fresh
# torch.zeros (freed when its last slice dies).  A chunk never spans a CUDA-graph capture boundary: the fill must be a node of the
# graph that uses it, so that every replay starts from zeros.
_ARENA = {}
_ARENA_BYTES = 2 << 20


def zeros(shape, dtype, device):
    n = 1
    for d in shape:
        n *= int(d)
    nbytes = ((n * torch.empty((), dtype=dtype).element_size() + 255) // 256) * 256
    device = torch.device(device)
    if device.type != "cuda" or nbytes > _ARENA_BYTES // 8 or os.environ.get("ICL_DISABLE_ZERO_ARENA") == "1":
        return torch.zeros(shape, dtype=dtype, device=device)
    capturing = torch.cuda.is_current_stream_capturing()
    key = (device.index if device.index is not None else torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
    a = _ARENA.get(key)
    if a is None or a[1] + nbytes > _ARENA_BYTES or a[2] != capturing:
        a = [torch.zeros((_ARENA_BYTES,), dtype=torch.uint8, device=device), 0, capturing]
        _ARENA[key] = a
    off = a[1]
    a[1] += nbytes
    return a[0][off:off + n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(tuple(shape))


def to_ndhwc(x):
    """[B,C,D,H,W] (any strides) -> contiguous [B,D,H,W,C] float32 (no copy if already channels-last)."""
    _require_cuda(x)
    if x.dtype != torch.float32:
        x = x.float()
    return x.permute(0, 2, 3, 4, 1).contiguous()


def to_ncdhw_view(x):
    """contiguous [B,D,H,W,C] -> [B,C,D,H,W] view."""
    return x.permute(0, 4, 1, 2, 3)


def empty_pk(B, C, D, H, W, device):
    return torch.empty((planes(), B, C // 8, D, H, W, 8), dtype=torch.bfloat16, device=device)


def pk_ok(C):
    return C % 16 == 0


# ------------------------------------------------------------------------------------------
# conv 3x3x3
# ------------------------------------------------------------------------------------------

def umma_ok(cins, cout, with_stats=True):
    """Shapes the tcgen05 implicit-GEMM kernel takes; everything else runs the fp32 CUDA-core kernel.
    ICL_DISABLE_UMMA=1 is a debugging knob that routes every conv to the CUDA-core kernel."""
    if os.environ.get("ICL_DISABLE_UMMA") == "1" or not tensor_cores():
        return False
    return all(c % 16 == 0 for c in cins) and cout % 16 == 0 and (cout <= 256 or not with_stats)


_MAX_CTAS = 0  # 0 = one CTA per SM; tests may lower it to exercise the persistent tile loop


def conv3d_umma(pks, cins, wp, bias, cout, B, D, H, W, stats=None, split=None, out=None):
    """pks: 1 or 2 PK tensors (virtual channel concat).  wp: packed weights from pack_w_umma (a ("walk", tensor) pair selects
    the plane-walk kernel).  Returns NDHWC output [B,D,H,W,cout], or a pair (y0 [..,split], y1 [..,cout-split]) when split
    is given (data gradient of a concatenated input)."""
    dev = pks[0].device
    if split is None:
        y0 = out if out is not None else torch.empty((B, D, H, W, cout), dtype=torch.float32, device=dev)
        y1, ld1, sp = None, 0, 0
    else:
        y0 = out if out is not None else torch.empty((B, D, H, W, split), dtype=torch.float32, device=dev)
        y1 = torch.empty((B, D, H, W, cout - split), dtype=torch.float32, device=dev)
        ld1, sp = cout - split, split
    pk1 = pks[1] if len(pks) > 1 else None
    c1 = cins[1] if len(cins) > 1 else 0
    walk = isinstance(wp, tuple)
    call("icl_conv3d_umma_walk_fwd" if walk else "icl_conv3d_umma_fwd", P(pks[0]), c_int(cins[0]), P(pk1), c_int(c1), P(wp[1] if walk else wp),
         P(bias), P(y0), c_int(y0.shape[-1]), P(y1), c_int(ld1), c_int(sp), P(stats), c_int(B), c_int(D), c_int(H), c_int(W), c_int(cout),
         c_int(planes()), c_int(_MAX_CTAS), gflop=2e-9 * 27 * sum(cins) * cout * B * D * H * W,
         tag="B%d r%d %s->%d" % (B, D, "+".join(map(str, cins)), cout))
    return y0 if split is None else (y0, y1)


def walk_ok(k_total, n, D):
    """The plane-walk kernel takes thin layers whose packed weights fit in shared memory (ICL_DISABLE_WALK=1 turns it off)."""
    if os.environ.get("ICL_DISABLE_WALK") == "1":
        return False
    return bool(_lib.lib().icl_conv3d_umma_walk_ok(k_total, n, D, planes()))


def pack_w_umma(w, dgrad, D=0):
    """torch conv weight [Cout,Cin,3,3,3] -> staged bf16 operand (conv3d_umma.cu), or ("walk", operand) in the
    plane-walk layout (conv3d_umma_walk.cu) when the layer qualifies at depth D."""
    cout, cin = w.shape[0], w.shape[1]
    n, k = (cin, cout) if dgrad else (cout, cin)
    wp = torch.empty(planes() * n * k * 27, dtype=torch.bfloat16, device=w.device)
    if D and walk_ok(k, n, D):
        call("icl_pack_w_walk", P(w), P(wp), c_int(cout), c_int(cin), c_int(1 if dgrad else 0), c_int(planes()))
        return ("walk", wp)
    nt = _lib.lib().icl_umma_ntile(n)
    call("icl_pack_w_umma", P(w), P(wp), c_int(cout), c_int(cin), c_int(1 if dgrad else 0), c_int(nt), c_int(planes()))
    return wp


# Packed tensor-core weight operands are cached per weight tensor: one pack per weight per optimizer step instead of one per conv
# call (labeled + unlabeled forward, the data-gradient operand in both backward passes; every sliding window at inference).
# An entry is valid while the owning tensor is alive, its version counter is unchanged (torch in-place updates, load_state_dict) and
# no icl_b200 optimizer step has run since (our SGD kernels update parameters through raw pointers: weights_changed()).
_WPACK = {}


def weights_changed():
    _WPACK.clear()


def pack_w_umma_cached(w, dgrad, D=0, owner=None):
    """pack_w_umma with the cache above.  `owner`: the tensor whose identity / version guards the entry (the nn.Parameter `w` was
    detached from); without it nothing is cached."""
    if owner is None or os.environ.get("ICL_DISABLE_WPACK_CACHE") == "1":
        return pack_w_umma(w, dgrad, D)
    import weakref
    key = (id(owner), bool(dgrad), int(D), planes())
    hit = _WPACK.get(key)
    if hit is not None and hit[0]() is owner and hit[1] == owner._version and hit[2] == w.data_ptr():
        return hit[3]
    wp = pack_w_umma(w, dgrad, D)
    _WPACK[key] = (weakref.ref(owner), owner._version, w.data_ptr(), wp)
    return wp


def repack_w_f32(w, dgrad):
    cout, cin = w.shape[0], w.shape[1]
    wp = torch.empty(27 * cin * cout, dtype=torch.float32, device=w.device)
    call("icl_repack_w_f32", P(w), P(wp), c_int(cout), c_int(cin), c_int(1 if dgrad else 0))
    return wp


def conv3d_direct(xs, cins, wp, bias, cout, B, D, H, W, stats=None):
    y = torch.empty((B, D, H, W, cout), dtype=torch.float32, device=xs[0].device)
    x1 = xs[1] if len(xs) > 1 else None
    c1 = cins[1] if len(cins) > 1 else 0
    call("icl_conv3d_direct_fwd", P(xs[0]), c_int(cins[0]), P(x1), c_int(c1), P(wp), P(bias), P(y), c_int(cout), c_int(0), P(stats),
         c_int(B), c_int(D), c_int(H), c_int(W), c_int(cout), gflop=2e-9 * 27 * sum(cins) * cout * B * D * H * W)
    return y


def stem_ok(cins, cout):
    return list(cins) == [1] and cout == 16


def conv3d_stem_fwd(x, w, bias, B, D, H, W, stats=None):
    y = torch.empty((B, D, H, W, 16), dtype=torch.float32, device=x.device)
    call("icl_conv3d_stem_fwd", P(x), P(w), P(bias), P(y), P(stats), c_int(B), c_int(D), c_int(H), c_int(W), c_int(16),
         gflop=2e-9 * 27 * 16 * B * D * H * W, mbytes=1e-6 * B * D * H * W * 68)
    return y


def conv3d_stem_wgrad(x, dy, B, D, H, W):
    dw = zeros((16, 1, 3, 3, 3), torch.float32, x.device)
    call("icl_conv3d_stem_wgrad", P(x), P(dy), P(dw), c_int(B), c_int(D), c_int(H), c_int(W), c_int(16),
         gflop=2e-9 * 27 * 16 * B * D * H * W, mbytes=1e-6 * B * D * H * W * 68)
    return dw


def conv3d_wgrad(xs, cins, dy, cout, B, D, H, W, want_bias=True):
    """Returns (dw [Cout, sum(cins), 3,3,3], dbias [Cout] or None)."""
    cin_total = sum(cins)
    dw = zeros((cout, cin_total, 3, 3, 3), torch.float32, dy.device)
    db = zeros((cout,), torch.float32, dy.device) if want_bias else None
    off = 0
    for i, (x, c) in enumerate(zip(xs, cins)):
        call("icl_conv3d_wgrad", P(x), c_int(c), P(dy), c_int(cout), P(dw), c_int(cin_total), c_int(off), P(db if i == 0 else None),
             c_int(B), c_int(D), c_int(H), c_int(W), gflop=2e-9 * 27 * c * cout * B * D * H * W, tag="B%d r%d %d->%d" % (B, D, c, cout))
        off += c
    return dw, db


def wgrad_ts():
    """True: the weight gradient runs conv3d_wgrad_ts.cu (dY operand in tensor memory); ICL_WGRAD=smem selects the older
    shared-memory-operand kernel (conv3d_wgrad_umma.cu, even depths only) for A/B measurements."""
    return os.environ.get("ICL_WGRAD", "ts") != "smem"


def wgrad_umma_ok(cins, cout, D=2):
    """Shapes the tcgen05 weight-gradient kernels take: channel multiples of 16 (the shared-memory-operand kernel also needs an
    EVEN depth, it walks two planes per step)."""
    if os.environ.get("ICL_DISABLE_UMMA") == "1" or not tensor_cores():
        return False
    return all(c % 16 == 0 for c in cins) and cout % 16 == 0 and (wgrad_ts() or D % 2 == 0)


def conv3d_wgrad_umma(x_pks, cins, dy_pk, cout, B, D, H, W, bx=0):
    """Tensor-core weight gradient from PK operands.  Returns dw [Cout, sum(cins), 3,3,3] (no bias gradient).
    bx: batch size the X operands were allocated with when only their first B samples take part (0 = B)."""
    cin_total = sum(cins)
    dev = dy_pk.device
    dw = torch.empty((cout, cin_total, 3, 3, 3), dtype=torch.float32, device=dev)
    off = 0
    ts = wgrad_ts()
    for pk, c in zip(x_pks, cins):
        if ts:
            n = _lib.lib().icl_conv3d_wgrad_ts_workspace(c, cout, B, D, H, W)
        else:
            n = _lib.lib().icl_conv3d_wgrad_umma_slots(c, cout, B, D, H, W) * 9 * 64 * 32
        if n <= 0:
            raise RuntimeError("conv3d_wgrad_umma: unsupported shape Cin=%d Cout=%d D=%d" % (c, cout, D))
        ws = torch.empty(n, dtype=torch.float32, device=dev)
        call("icl_conv3d_wgrad_ts" if ts else "icl_conv3d_wgrad_umma", P(pk), c_int(c), P(dy_pk), c_int(cout), P(dw), c_int(cin_total), c_int(off),
             P(ws), c_int(B), c_int(D), c_int(H), c_int(W), c_int(planes()), c_int(0), c_int(bx),
             gflop=2e-9 * 27 * c * cout * B * D * H * W, tag="B%d r%d %d->%d" % (B, D, c, cout))
        off += c
    return dw


# ------------------------------------------------------------------------------------------
# InstanceNorm + ReLU, pool, upsample, dropout
# ------------------------------------------------------------------------------------------

def instnorm_finalize(stats, B, C, S, eps=1e-5):
    mr = torch.empty((B, C, 2), dtype=torch.float32, device=stats.device)
    call("icl_instnorm_finalize", P(stats), P(mr), c_int(B), c_int(C), c_ll(S), c_f(eps))
    return mr


def instnorm_relu_fwd(y, mr, want_pk, want_f32=True):
    """want_f32=False (needs want_pk): only the split-bf16 operand is written — the activation between the two convolutions of a
    UnetConv3 block is read by nothing but the next tensor-core convolution and its weight gradient."""
    B, D, H, W, C = y.shape
    a = torch.empty_like(y) if (want_f32 or not want_pk) else None
    pk = empty_pk(B, C, D, H, W, y.device) if want_pk else None
    call("icl_instnorm_relu_fwd", P(y), P(mr), P(a), P(pk), c_int(1 if planes() == 2 else 0), c_int(B), c_int(C), c_ll(D * H * W),
         mbytes=1e-6 * y.numel() * (4 + (4 if a is not None else 0) + (2 * planes() if want_pk else 0)))
    return a, pk


def instnorm_relu_bwd(dA, y, mr, want_pk, want_dbias=False, want_f32=True):
    """Returns (dY or None, dY_pk or None[, dbias]) — dbias = sum of dY over samples and voxels (the conv-bias gradient)."""
    B, D, H, W, C = y.shape
    red = zeros((B, C, 2), torch.float64, y.device)
    dY = torch.empty_like(y) if want_f32 else None
    pk = empty_pk(B, C, D, H, W, y.device) if want_pk else None
    db = zeros((C,), torch.float32, y.device) if want_dbias else None
    call("icl_instnorm_relu_bwd", P(dA), P(y), P(mr), P(red), P(dY), P(pk), c_int(1 if planes() == 2 else 0), P(db), c_int(B), c_int(C),
         c_ll(D * H * W), tag="B%d r%d C%d" % (B, D, C), mbytes=1e-6 * y.numel() * (16 + (4 if want_f32 else 0) + (2 * planes() if want_pk else 0)))
    return (dY, pk, db) if want_dbias else (dY, pk)


def pack_pk(x):
    B, D, H, W, C = x.shape
    pk = empty_pk(B, C, D, H, W, x.device)
    call("icl_pack_pk", P(x), P(pk), c_int(1 if planes() == 2 else 0), c_int(B), c_int(C), c_ll(D * H * W))
    return pk


def maxpool_fwd(a, want_pk, want_f32=True):
    B, D, H, W, C = a.shape
    out = torch.empty((B, D // 2, H // 2, W // 2, C), dtype=torch.float32, device=a.device) if want_f32 else None
    idx = torch.empty((B, D // 2, H // 2, W // 2, C), dtype=torch.uint8, device=a.device)
    pk = empty_pk(B, C, D // 2, H // 2, W // 2, a.device) if want_pk else None
    call("icl_maxpool3d_fwd", P(a), P(out), P(idx), P(pk), c_int(1 if planes() == 2 else 0), c_int(B), c_int(C), c_int(D), c_int(H), c_int(W))
    return out, idx, pk


def maxpool_bwd(dout, idx, dx, accumulate):
    B, D, H, W, C = dx.shape
    call("icl_maxpool3d_bwd", P(dout), P(idx), P(dx), c_int(1 if accumulate else 0), c_int(B), c_int(C), c_int(D), c_int(H), c_int(W))


def upsample2x_fwd(x, want_pk, want_f32=True):
    B, d, h, w, C = x.shape
    out = torch.empty((B, 2 * d, 2 * h, 2 * w, C), dtype=torch.float32, device=x.device) if want_f32 else None
    pk = empty_pk(B, C, 2 * d, 2 * h, 2 * w, x.device) if want_pk else None
    call("icl_upsample2x_fwd", P(x), P(out), P(pk), c_int(1 if planes() == 2 else 0), c_int(B), c_int(C), c_int(d), c_int(h), c_int(w),
         mbytes=1e-6 * x.numel() * (4 + 8 * ((4 if want_f32 else 0) + (2 * planes() if want_pk else 0))))
    return out, pk


def upsample2x_bwd(dout, c_off, C, dx, accumulate):
    """dout: NDHWC fine-res gradient with dout.shape[-1] channels, this op owns [c_off, c_off+C)."""
    B, d, h, w, _ = dx.shape
    call("icl_upsample2x_bwd", P(dout), c_int(dout.shape[-1]), c_int(c_off), P(dx), c_int(1 if accumulate else 0), c_int(B), c_int(C),
         c_int(d), c_int(h), c_int(w))


def dropout(x, p, mask=None, seed=0):
    """seed: python int, or a 1-element int64 CUDA tensor (read by the kernel: no host sync, CUDA-graph capturable)."""
    out = torch.empty_like(x)
    if isinstance(seed, torch.Tensor):
        call("icl_dropout", P(x), P(out), P(mask), c_ull(0), P(seed), c_f(p), c_ll(x.numel()))
    else:
        call("icl_dropout", P(x), P(out), P(mask), c_ull(int(seed) & 0xFFFFFFFFFFFFFFFF), P(None), c_f(p), c_ll(x.numel()))
    return out


# ------------------------------------------------------------------------------------------
# GEMM family
# ------------------------------------------------------------------------------------------

def sgemm(M, N, K, A, sam, sak, Bm, sbk, sbn, C, scm, scn, bias=None, bias_mode=0, act=0, accumulate=False, pre=None, batch=1, sA=0, sB=0,
          sC=0):
    call("icl_sgemm", c_int(M), c_int(N), c_int(K), P(A), c_ll(sam), c_ll(sak), c_ll(sA), P(Bm), c_ll(sbk), c_ll(sbn), c_ll(sB), P(C),
         c_ll(scm), c_ll(scn), c_ll(sC), c_int(batch), P(bias), c_int(bias_mode), c_int(act), c_int(1 if accumulate else 0), P(pre),
         gflop=2e-9 * M * N * K * batch, mbytes=4e-6 * batch * (M * K + K * N + M * N), tag="%dx%dx%d b%d" % (M, N, K, batch))


BIGW_MIN_NUMEL = 1 << 21  # weights at least this large (mlp2 at the 24^3 and 12^3 levels) stream through the tcgen05 kernels (bigw_umma.cu)


def bigw_ok(M, N, K):
    """Linear shapes the tcgen05 weight-streaming kernels take: a big weight, few rows (ICL_DISABLE_BIGW=1 turns them off)."""
    if os.environ.get("ICL_DISABLE_BIGW") == "1" or not tensor_cores():
        return False
    return N * K >= BIGW_MIN_NUMEL and M <= 512 and K % 4 == 0 and N % 4 == 0


def tok_linear_ok(M, N, K):
    """Token-major Linears the tcgen05 kernel takes (many tokens against a small weight: Swin-UNet qkv / proj / mlp, the ICL-head
    token projections); ICL_DISABLE_TOKLIN=1 routes them back to the fp32 CUDA-core GEMM."""
    if os.environ.get("ICL_DISABLE_TOKLIN") == "1" or not tensor_cores():
        return False
    # below ~0.8 GFLOP the pack + GEMM (+ finish) launches cost more than the fp32 CUDA-core GEMM they replace (measured on the ICL-head
    # token projections, 27 648 x 64 x 64: the replayed step got 0.3 ms slower); ICL_TOKLIN_MIN_FLOP overrides (tests use 0)
    min_flop = float(os.environ.get("ICL_TOKLIN_MIN_FLOP", "0.8e9"))
    return M >= 512 and N >= 16 and K >= 32 and K % 4 == 0 and N % 4 == 0 and 2.0 * M * N * K >= min_flop


def _tok_ws(N, K, device):
    return torch.empty((int(_lib.lib().icl_tok_linear_workspace(N, K)),), dtype=torch.uint8, device=device)


def _bigw_ws(rows, M, K, device):
    return torch.empty((int(_lib.lib().icl_bigw_workspace(rows, M, K)),), dtype=torch.uint8, device=device)


def linear_fwd(x2d, w, b, act=0, want_pre=False):
    """y = act(x2d @ w.T + b); x2d [M,K] contiguous, w [N,K]."""
    M, K = x2d.shape
    N = w.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=x2d.device)
    pre = torch.empty_like(y) if want_pre else None
    if bigw_ok(M, N, K):
        ws = _bigw_ws(M, N, K, x2d.device)
        call("icl_bigw_linear_fwd", c_int(M), c_int(N), c_int(K), P(x2d), P(w), P(b), P(y), P(pre), c_int(act), P(ws),
             mbytes=4e-6 * (N * K + M * K + M * N), gflop=2e-9 * M * N * K, tag="%dx%dx%d" % (M, N, K))
    elif tok_linear_ok(M, N, K):
        call("icl_tok_linear_fwd", c_int(M), c_int(N), c_int(K), P(x2d), P(w), P(b), P(y), P(pre), c_int(act), P(_tok_ws(N, K, x2d.device)),
             mbytes=4e-6 * (N * K + M * K + M * N * (2 if want_pre else 1)), gflop=2e-9 * M * N * K, tag="%dx%dx%d" % (M, N, K))
    elif M <= 64 and K >= 1024 and K % 4 == 0:
        call("icl_skinny_linear_fwd", c_int(M), c_int(N), c_int(K), P(x2d), P(w), P(b), P(y), P(pre), c_int(act),
             mbytes=4e-6 * (N * K + M * K + M * N), gflop=2e-9 * M * N * K, tag="%dx%dx%d" % (M, N, K))
    else:
        sgemm(M, N, K, x2d, K, 1, w, 1, K, y, N, 1, bias=b, bias_mode=1 if b is not None else 0, act=act, pre=pre)
    return y, pre


def linear_dgrad(dy2d, w):
    M, N = dy2d.shape
    K = w.shape[1]
    if bigw_ok(M, N, K):
        dx = torch.empty((M, K), dtype=torch.float32, device=dy2d.device)
        ws = _bigw_ws(M, K, N, dy2d.device)
        call("icl_bigw_linear_dgrad", c_int(M), c_int(N), c_int(K), P(dy2d), P(w), P(dx), P(ws), mbytes=4e-6 * (N * K + M * K + M * N),
             gflop=2e-9 * M * N * K, tag="%dx%dx%d" % (M, N, K))
        return dx
    if tok_linear_ok(M, K, N):
        dx = torch.empty((M, K), dtype=torch.float32, device=dy2d.device)
        call("icl_tok_linear_dgrad", c_int(M), c_int(N), c_int(K), P(dy2d), P(w), P(dx), P(_tok_ws(K, N, dy2d.device)),
             mbytes=4e-6 * (N * K + M * K + M * N), gflop=2e-9 * M * N * K, tag="%dx%dx%d" % (M, N, K))
        return dx
    if M <= 64 and N >= 1024 and K % 4 == 0:
        dx = zeros((M, K), torch.float32, dy2d.device)
        call("icl_skinny_linear_dgrad", c_int(M), c_int(N), c_int(K), P(dy2d), P(w), P(dx), mbytes=4e-6 * (N * K + M * K + M * N),
             gflop=2e-9 * M * N * K, tag="%dx%dx%d" % (M, N, K))
    else:
        dx = torch.empty((M, K), dtype=torch.float32, device=dy2d.device)
        sgemm(M, K, N, dy2d, N, 1, w, K, 1, dx, K, 1)
    return dx


def linear_wgrad(dy2d, x2d, want_bias=True):
    """dW [N,K] = dy^T x ; db [N] = colsum(dy)."""
    M, N = dy2d.shape
    K = x2d.shape[1]
    dW = torch.empty((N, K), dtype=torch.float32, device=dy2d.device)
    db = torch.empty((N,), dtype=torch.float32, device=dy2d.device) if want_bias else None
    if tok_linear_ok(M, N, K) and M >= 2048 and K % 4 == 0:
        ws = torch.empty((int(_lib.lib().icl_tok_linear_wgrad_workspace(M, N, K)),), dtype=torch.uint8, device=dy2d.device)
        call("icl_tok_linear_wgrad", c_int(M), c_int(N), c_int(K), P(dy2d), P(x2d), P(dW), P(ws), mbytes=4e-6 * (N * K + M * K + M * N),
             gflop=2e-9 * M * N * K, tag="%dx%dx%d" % (M, N, K))
        if want_bias:
            call("icl_colsum", P(dy2d), P(db), c_ll(M), c_int(N), c_int(0), tag="%dx%d" % (M, N))
    elif M <= 64 and N * K >= (1 << 20):
        call("icl_outer_wgrad", c_int(M), c_int(N), c_int(K), P(dy2d), P(x2d), P(dW), P(db), c_int(0),
             mbytes=4e-6 * (N * K + M * K + M * N), gflop=2e-9 * M * N * K)
    else:
        sgemm(N, K, M, dy2d, 1, N, x2d, K, 1, dW, K, 1)
        if want_bias:
            call("icl_colsum", P(dy2d), P(db), c_ll(M), c_int(N), c_int(0), tag="%dx%d" % (M, N))
    return dW, db


def outer_wgrad_acc(dy2d, x2d, dW):
    """dW (+)= dy^T x for <= 64 rows (factor-exchange data parallelism, icl_b200/parallel.py)."""
    M, N = dy2d.shape
    K = x2d.shape[1]
    acc = dW is not None
    if dW is None:
        dW = torch.empty((N, K), dtype=torch.float32, device=dy2d.device)
    call("icl_outer_wgrad", c_int(M), c_int(N), c_int(K), P(dy2d), P(x2d), P(dW), P(None), c_int(1 if acc else 0),
         mbytes=4e-6 * N * K * (2 if acc else 1))
    return dW


def gelu_bwd(dy, pre):
    dx = torch.empty_like(pre)
    call("icl_gelu_bwd", P(dy), P(pre), P(dx), c_ll(pre.numel()), mbytes=12e-6 * pre.numel())
    return dx


def axpby(x, y, alpha, beta):
    call("icl_axpby", P(x), P(y), c_f(alpha), c_f(beta), c_ll(x.numel()))


def row_combine(a, sa, b, sb, rows):
    out = torch.empty_like(a)
    call("icl_row_combine", P(a), P(sa), P(b), P(sb), P(out), c_ll(rows), c_ll(a.numel() // rows), mbytes=4e-6 * a.numel() * (3 if b is not None else 2))
    return out


def sgd_factored(p, m, factors, lr, mu, wd, max_ctas=0):
    """Momentum-SGD update of a huge 2-D weight from its rank-R gradient factors [(dy [r,N], x [r,K], scale), ...]:
    g = sum dy^T x (+ wd p), m = mu m + g, p -= lr m.  tcgen05 path when the shape allows, CUDA-core kernel otherwise."""
    N, K = p.shape
    R = sum(f[0].shape[0] for f in factors)
    if tensor_cores() and N % 8 == 0 and K % 8 == 0 and os.environ.get("ICL_DISABLE_BIGW") != "1":
        ws = torch.empty((int(_lib.lib().icl_sgd_factored_workspace(R, N, K)),), dtype=torch.uint8, device=p.device)
        r0 = 0
        for dy, x, scale in factors:
            call("icl_sgd_factored_pack", P(dy), P(x), c_int(dy.shape[0]), c_int(r0), c_int(R), c_int(N), c_int(K), c_f(scale), P(ws))
            r0 += dy.shape[0]
        call("icl_sgd_factored_apply", c_int(R), c_int(N), c_int(K), P(ws), P(p), P(m), P(lr), c_f(mu), c_f(wd), c_int(max_ctas or _MAX_CTAS),
             mbytes=16e-6 * p.numel(), gflop=2e-9 * R * N * K, tag="R%d %dx%d" % (R, N, K))
    else:
        dy = torch.cat([f[0] * f[2] if f[2] != 1.0 else f[0] for f in factors], 0).contiguous()
        x = torch.cat([f[1] for f in factors], 0).contiguous()
        call("icl_sgd_factored", c_int(R), c_int(N), c_int(K), P(dy), P(x), P(p), P(m), P(lr), c_f(mu), c_f(wd),
             mbytes=16e-6 * p.numel(), tag="R%d %dx%d" % (R, N, K))
