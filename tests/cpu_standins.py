"""Plain-torch stand-ins for the icl_b200.functional kernels (TEST INFRASTRUCTURE, CPU only).

The container that runs `pytest -m "not gpu"` has no GPU, so the HOST LOGIC of the network mirrors (module wiring, token
re-arrangements, parameter naming, which parameters receive gradients) is exercised by swapping every kernel-backed function
of icl_b200.functional for a differentiable torch expression with the same signature.  Nothing in the package imports this
file; the product path has no CPU fallback (icl_b200.ops._require_cuda raises on CPU tensors)."""
import torch
import torch.nn.functional as F

from oracle import restate_swin as RS


def linear(x, w, b=None, act=0):
    y = F.linear(x, w, b)
    return F.gelu(y) if act else y


def layer_norm(x, w, b, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def add_scaled(a, b, r=None):
    if r is None:
        return a + b
    return a + b * r.reshape((-1,) + (1,) * (b.dim() - 1))


def batch_mean(x):
    return x.mean(dim=0, keepdim=True)


def proxy_attention(ql, kv, num_heads, want_xv=True):
    B, K, C = ql.shape
    N = kv.shape[1]
    hd = C // num_heads
    q = ql.reshape(B, num_heads, K, hd)
    k, v = kv.reshape(B, N, 2, num_heads, hd).permute(2, 0, 3, 1, 4)
    a = (q @ k.transpose(-2, -1)) * hd ** -0.5
    xv = (a.softmax(-1) @ v).reshape(B, K, C) if want_xv else None
    return xv, a.permute(0, 2, 1, 3)


def dwconv3d(x, w):
    return F.conv3d(x, w, None, 1, 1, 1, groups=x.shape[1])


def bn_relu(x, bn, training):
    if training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1
    return F.relu(F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, training, bn.momentum, bn.eps))


def planar_pointwise(x, w, b=None):
    NB, CI = x.shape[0], x.shape[1]
    y = torch.einsum("oc,ncs->nos", w.reshape(w.shape[0], CI), x.reshape(NB, CI, -1))
    if b is not None:
        y = y + b[None, :, None]
    return y.reshape((NB, w.shape[0]) + tuple(x.shape[2:]))


def window_attention(qkv, table, H, W, num_heads, window_size, shift):
    return RS.window_attention_tokens(qkv, table, H, W, num_heads, window_size, shift)


ALL = dict(linear=linear, layer_norm=layer_norm, add_scaled=add_scaled, batch_mean=batch_mean, proxy_attention=proxy_attention,
           dwconv3d=dwconv3d, bn_relu=bn_relu, planar_pointwise=planar_pointwise, window_attention=window_attention)


def install(monkeypatch):
    import icl_b200.functional as Fn
    for k, v in ALL.items():
        monkeypatch.setattr(Fn, k, v)
