"""CPU, reference-free: the 2D oracle restatement (oracle/restate2d.py) against the committed fixtures that the unmodified
reference produced (tests/golden/unet2d_k4_64, losses2d_k4, step_cfg1; generator oracle/make_golden.py).  Parameters are rebuilt
from the name/shape fixture (state_keys.json) with the same synthetic generator the fixtures used."""
import json
import os
from collections import OrderedDict

import pytest
import torch

from helpers import GOLDEN, assert_close, check_summary, golden
from oracle import restate as R
from oracle import restate2d as R2
from oracle import synth

TOL = 2e-5


def _params(key, seed):
    keys = json.load(open(os.path.join(GOLDEN, "state_keys.json")))
    shapes = OrderedDict((k, tuple(s)) for k, s in keys[key])
    return R.make_params(synth.synth_state_dict(shapes, seed)), keys[key + "_params"]


def test_unet2d_golden_oracle():
    g = golden("unet2d_k4_64")
    K, size, seed, B = [int(v) for v in g["meta"]]
    P, names = _params("unet2d_k4", seed)
    x = synth.synth_volume((B, 1, size, size), seed + 1)
    y = synth.synth_labels((B, size, size), K, seed + 2)
    logits = R2.unet2d_forward(P, x)
    assert_close(logits.detach(), g["logits"], TOL, "2D logits")
    loss = R.ce_loss(logits, y) + R.dice_loss(logits, y.unsqueeze(1), K, softmax=True)
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    loss.backward()
    for k in names:
        check_summary(P[k].grad, g["gsum/" + k], g["gval/" + k], 2e-4, k, abs_floor=1e-6)
    for k, v in P.items():
        if "running" in k:
            assert_close(v, g["stat/" + k], 1e-5, k, abs_floor=1e-7)


def test_losses2d_golden_oracle():
    g = golden("losses2d_k4")
    K, S, B = [int(v) for v in g["meta"]]
    labels = synth.synth_labels((B, S, S), K, 51)
    out_lab = synth.synth_volume((B, K, S, S), 52).requires_grad_(True)
    out_unlab = synth.synth_volume((B, K, S, S), 53)
    mk = lambda s: [synth.synth_volume((B, K, r, r), s + i).mul_(2.0) for i, r in enumerate((8, 16, 32))]
    fms = [t.requires_grad_(True) for t in mk(60)]
    fms2 = [t.requires_grad_(True) for t in mk(70)]
    fms3 = mk(80)
    L = R2.icl_losses_2d((out_lab, out_unlab, fms, fms2, fms3), labels, K)
    for nm in ("ce", "dice", "aux", "pse", "cons", "total"):
        assert abs(float(L[nm]) - float(g[nm])) <= 1e-5 * max(1.0, abs(float(g[nm]))), nm
    L["total"].backward()
    assert_close(out_lab.grad, g["dout"], 1e-5, "dout")
    for i in range(3):
        assert_close(fms[i].grad, g["daux%d" % i], 1e-5, "daux%d" % i)
        assert_close(fms2[i].grad, g["dpse%d" % i], 1e-5, "dpse%d" % i)


def test_step_cfg1_golden_oracle():
    """BASELINE config 1 (the reference's own CPU-runnable case): UNet_icl(1, 4), 12 + 12 slices of 1x256x256."""
    g = golden("step_cfg1")
    K = int(g["K"])
    P, names = _params("unet_icl_k4", 1337)
    x = synth.synth_volume((24, 1, 256, 256), 1338)
    y = synth.synth_labels((24, 256, 256), K, 1339)
    o = R2.unet_icl_forward(P, x[:12], x[12:])
    L = R2.icl_losses_2d(o, y[:12], K)
    for nm in ("ce", "dice", "aux", "pse", "cons", "total"):
        assert abs(float(L[nm]) - float(g[nm])) <= 1e-4 * max(abs(float(g[nm])), 1e-3), (nm, float(L[nm]), float(g[nm]))
    for nm, t in (("out_lab", o[0]), ("out_unlab", o[1])):
        check_summary(t, g[nm + "_sum"], g[nm + "_val"], 5e-5, nm, n=4096)
    for j, nm in ((2, "maps_lab"), (3, "maps_unlab"), (4, "maps_consis")):
        for i in range(3):
            check_summary(o[j][i], g["%s%d_sum" % (nm, i)], g["%s%d_val" % (nm, i)], 1e-4, "%s%d" % (nm, i), n=4096)
    L["total"].backward()
    none = set(str(s) for s in g["grad_none"])
    for k in names:
        if k in none:
            assert P[k].grad is None, k
        else:
            check_summary(P[k].grad, g["gsum/" + k], g["gval/" + k], 1e-3, k, abs_floor=5e-6)  # floor: the exactly-zero conv-bias gradients
