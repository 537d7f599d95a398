"""Shared comparison helpers for the oracle and GPU parity tests."""
import os

import numpy as np
import torch

from oracle.make_golden import sample_idx, unpack_labels  # same seeded sample positions / label packing as the fixtures

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def rel_l2(a, b):
    a = torch.as_tensor(np.asarray(a)).double().reshape(-1)
    b = torch.as_tensor(np.asarray(b)).double().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def assert_close(a, b, rel, what="", abs_floor=0.0):
    """rel-L2 of the whole tensor <= rel, or (for mathematically-zero tensors) l2 <= abs_floor."""
    a = torch.as_tensor(np.asarray(a)).double().reshape(-1)
    b = torch.as_tensor(np.asarray(b)).double().reshape(-1)
    err = (a - b).norm().item()
    ref = b.norm().item()
    assert err <= rel * ref + abs_floor, "%s: |a-b|=%.3e |b|=%.3e rel=%.3e (tol %.1e)" % (what, err, ref, err / max(ref, 1e-30), rel)


def check_summary(t, gsum, gval, rel, what, n=64, abs_floor=1e-7):
    """Compare a tensor against a fixture summary ([sum, l2], sampled values)."""
    t = t.detach().double().cpu().reshape(-1)
    idx = sample_idx(t.numel(), n)
    l2 = float(gsum[1])
    assert abs(t.norm().item() - l2) <= rel * l2 + abs_floor, "%s: l2 %.6e vs %.6e" % (what, t.norm().item(), l2)
    got = t[idx]
    want = torch.from_numpy(np.asarray(gval)).double()
    # sampled values: error relative to the tensor's RMS (individual entries may be ~0)
    rms = l2 / max(t.numel(), 1) ** 0.5
    err = (got - want).abs().max().item()
    assert err <= 30 * rel * rms + rel * want.abs().max().item() + abs_floor, \
        "%s: sampled max err %.3e (rms %.3e)" % (what, err, rms)


def assert_argmax_agrees(logits, packed_bits, what, min_agree=0.999):
    """Per-voxel agreement of argmax(logits, dim 1) with the reference's label map stored as packed bit planes in the fixture
    (north star: label maps agree on at least 99.9 % of voxels)."""
    got = logits.detach().argmax(1).reshape(-1).cpu().numpy()
    want = unpack_labels(np.asarray(packed_bits), got.size)
    agree = float((got == want).mean())
    assert agree >= min_agree, "%s: argmax agrees on %.5f of %d voxels (need %.4f)" % (what, agree, got.size, min_agree)
    return agree
