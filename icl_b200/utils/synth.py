"""Deterministic synthetic parameters and inputs shared by the golden-fixture generator,
the oracle tests and the GPU parity tests (test infrastructure; numpy + torch only, so it
travels to the GPU box).

Why not torch.manual_seed + the reference constructors: constructor init consumes the global
torch RNG in construction order, which a different module tree cannot
reproduce, and 785 M parameters cannot be committed.  Instead every tensor of a state_dict is
filled from its own PCG64 stream keyed by (seed, position in state_dict order), so the same
values can be regenerated anywhere from the parameter *shapes* alone.  Scales follow the
reference init (kaiming fan_in for convs, networks_other.py:40-49; torch default uniform for
Linear) so the numerics are representative; LayerNorm/BatchNorm affine terms and guided_Q are
perturbed away from 1/0 so tests exercise them (guided_Q is zero at init).
"""
from collections import OrderedDict

import numpy as np
import torch


def _rng(seed, idx):
    return np.random.Generator(np.random.PCG64([int(seed), int(idx)]))


def _uniform(rng, shape, scale):
    n = int(np.prod(shape)) if len(shape) else 1
    a = rng.random(n, dtype=np.float32)
    a *= np.float32(2.0 * scale)
    a -= np.float32(scale)
    return a.reshape(shape)


def synth_tensor(name, shape, seed, idx, dtype=torch.float32):
    shape = tuple(int(s) for s in shape)
    rng = _rng(seed, idx)
    if name.endswith("num_batches_tracked"):
        return torch.zeros(shape, dtype=torch.int64)
    if name.endswith("running_mean"):
        return torch.from_numpy(_uniform(rng, shape, 0.1))
    if name.endswith("running_var"):
        return torch.from_numpy(1.0 + _uniform(rng, shape, 0.1))
    if name.endswith("guided_Q") or name.endswith("guide_Q"):
        a = rng.standard_normal(int(np.prod(shape)), dtype=np.float32).reshape(shape)
        return torch.from_numpy(a * np.float32(0.5))
    if len(shape) >= 3:  # conv weight [Cout, Cin/groups, k...]
        fan_in = int(np.prod(shape[1:]))
        a = rng.standard_normal(int(np.prod(shape)), dtype=np.float32).reshape(shape)
        a *= np.float32(np.sqrt(2.0 / fan_in))
        return torch.from_numpy(a)
    if len(shape) == 2:  # Linear weight [out, in]
        return torch.from_numpy(_uniform(rng, shape, 1.0 / np.sqrt(shape[1])))
    # 1-D: norm weight (≈1) or any bias (small)
    if name.endswith(".weight"):
        return torch.from_numpy(1.0 + _uniform(rng, shape, 0.1))
    return torch.from_numpy(_uniform(rng, shape, 0.1))


STRUCTURAL = ("relative_position_index", "attn_mask")  # Swin index / mask buffers: functions of the geometry, never synthesised


def synth_state_dict(shapes, seed):
    """shapes: OrderedDict name -> shape (e.g. from model.state_dict()).  Returns name -> tensor."""
    out = OrderedDict()
    for idx, (name, shape) in enumerate(shapes.items()):
        if not name.endswith(STRUCTURAL):
            out[name] = synth_tensor(name, shape, seed, idx)
    return out


def shapes_of(model):
    return OrderedDict((k, tuple(v.shape)) for k, v in model.state_dict().items())


def load_synth(model, seed):
    """Overwrite every entry of model.state_dict() in place with synthetic values."""
    sd = model.state_dict()
    with torch.no_grad():
        for idx, (name, t) in enumerate(sd.items()):
            if name.endswith(STRUCTURAL):
                continue
            v = synth_tensor(name, t.shape, seed, idx)
            t.copy_(v.to(t.dtype))
    return model


def synth_volume(shape, seed, kind="randn"):
    rng = _rng(seed, 10_000_019)
    n = int(np.prod(shape))
    if kind == "randn":
        a = rng.standard_normal(n, dtype=np.float32)
    else:
        a = rng.random(n, dtype=np.float32)
    return torch.from_numpy(a.reshape(shape))


def synth_labels(shape, num_classes, seed):
    rng = _rng(seed, 10_000_079)
    a = rng.integers(0, num_classes, size=int(np.prod(shape)), dtype=np.int64)
    return torch.from_numpy(a.reshape(shape))


def synth_blobs(shape, num_classes, seed):
    """Spatially coherent labels (coarse random grid upsampled by nearest) — closer to real
    segmentation masks than iid noise; used for Dice/metric tests."""
    rng = _rng(seed, 10_000_103)
    coarse = [max(1, s // 8) for s in shape[-3:]]
    lead = tuple(shape[:-3])
    a = rng.integers(0, num_classes, size=lead + tuple(coarse), dtype=np.int64)
    for ax, s in zip((-3, -2, -1), shape[-3:]):
        rep = -(-s // a.shape[ax])
        a = np.repeat(a, rep, axis=ax)
        a = np.take(a, np.arange(s), axis=ax)
    return torch.from_numpy(np.ascontiguousarray(a))
