"""CPU: the oracle restatement (oracle/restate.py) against the fixtures produced by the
unmodified reference (tests/golden, generator oracle/make_golden.py)."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import restate as R
from oracle import synth
from oracle.make_golden import MINI
from helpers import assert_close, check_summary, golden

TOL = 2e-5  # fp32 re-association noise between two CPU evaluations of the same math


def backbone_shapes(fs, K, cin):
    f = [int(x / fs) for x in (64, 128, 256, 512, 1024)]
    sh = OrderedDict()
    def block(name, ci, co):
        sh[name + ".conv1.0.weight"] = (co, ci, 3, 3, 3); sh[name + ".conv1.0.bias"] = (co,)
        sh[name + ".conv2.0.weight"] = (co, co, 3, 3, 3); sh[name + ".conv2.0.bias"] = (co,)
    block("conv1", cin, f[0]); block("conv2", f[0], f[1]); block("conv3", f[1], f[2]); block("conv4", f[2], f[3])
    block("center", f[3], f[4])
    block("up_concat4.conv", f[4] + f[3], f[3]); block("up_concat3.conv", f[3] + f[2], f[2])
    block("up_concat2.conv", f[2] + f[1], f[1]); block("up_concat1.conv", f[1] + f[0], f[0])
    sh["final.weight"] = (K, f[0], 1, 1, 1); sh["final.bias"] = (K,)
    return sh


@pytest.mark.parametrize("name", ["unet3d_fs4_k2_c1_32", "unet3d_fs16_k3_c2_32"])
def test_backbone_golden(name):
    g = golden(name)
    fs, K, cin, size, seed, B = [int(v) for v in g["meta"]]
    P = R.make_params(synth.synth_state_dict(backbone_shapes(fs, K, cin), seed))
    x = synth.synth_volume((B, cin) + (size,) * 3, seed + 1)
    y = synth.synth_labels((B,) + (size,) * 3, K, seed + 2)
    logits = R.unet_3d_forward(P, x)
    assert_close(logits.detach(), g["logits"], TOL, "logits")
    loss = R.ce_loss(logits, y) + R.dice_loss(torch.softmax(logits, 1), y.unsqueeze(1), K)
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    loss.backward()
    for k, p in P.items():
        check_summary(p.grad, g["gsum/" + k], g["gval/" + k], 2e-4, k)


def mini_head_shapes():
    c = MINI
    sh = OrderedDict()
    K = c["K"]
    groups = ["proj_layers", "norm_layers", "class_decoders", "attn_convs0", "attn_convs1", "query_convs"]
    per = {g: [] for g in groups}
    for i, (C, r, H) in enumerate(zip(c["in_chans"], c["res"], c["heads"])):
        N = r ** 3
        per["proj_layers"] += [("%d.weight" % i, (C, C, 1, 1, 1)), ("%d.bias" % i, (C,))]
        per["norm_layers"] += [("%d.weight" % i, (C,)), ("%d.bias" % i, (C,))]
        cd = "%d." % i
        for nm, s in (("norm1", C), ("norm1_query", C)):
            per["class_decoders"] += [(cd + nm + ".weight", (s,)), (cd + nm + ".bias", (s,))]
        per["class_decoders"] += [(cd + "attn.fc_q.weight", (C, C)), (cd + "attn.fc_q.bias", (C,)),
                                  (cd + "attn.fc_kv.weight", (2 * C, C)), (cd + "attn.fc_kv.bias", (2 * C,)),
                                  (cd + "attn.proj.weight", (C, C)), (cd + "attn.proj.bias", (C,)),
                                  (cd + "norm2.weight", (C,)), (cd + "norm2.bias", (C,)),
                                  (cd + "mlp.fc1.weight", (4 * C, C)), (cd + "mlp.fc1.bias", (4 * C,)),
                                  (cd + "mlp.fc2.weight", (C, 4 * C)), (cd + "mlp.fc2.bias", (C,)),
                                  (cd + "norm3.weight", (N,)), (cd + "norm3.bias", (N,)),
                                  (cd + "mlp2.fc1.weight", (N, N)), (cd + "mlp2.fc1.bias", (N,)),
                                  (cd + "mlp2.fc2.weight", (N, N)), (cd + "mlp2.fc2.bias", (N,))]
        b = "%d.block." % i
        per["attn_convs0"] += [(b + "depthwise.weight", (H, 1, 3, 3, 3))]
        for bn in ("bn_depth", "bn_point"):
            if bn == "bn_point":
                per["attn_convs0"] += [(b + "pointwise.weight", (H, H, 1, 1, 1))]
            per["attn_convs0"] += [(b + bn + ".weight", (H,)), (b + bn + ".bias", (H,)),
                                   (b + bn + ".running_mean", (H,)), (b + bn + ".running_var", (H,)),
                                   (b + bn + ".num_batches_tracked", ())]
        per["attn_convs1"] += [("%d.weight" % i, (1, H, 1, 1, 1)), ("%d.bias" % i, (1,))]
        per["query_convs"] += [("%d.weight" % i, (C // 2, C, 1)), ("%d.bias" % i, (C // 2,))]
    sh["guided_Q"] = (1, K, c["in_chans"][0])
    for g in groups:
        for nm, s in per[g]:
            sh[g + "." + nm] = s
    return sh


def test_mini_head_shapes_match_reference_order():
    """state_dict order matters for synth (tensor idx = position)."""
    pytest.importorskip("torch")
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference not present")
    ns = ref_import.load()
    c = MINI
    ic = ns.InherentConsistent(in_chans=c["in_chans"], depths=(2, 2, 2), patch_size=(2, 2, 2),
                               input_resolution=c["res"], num_classes=c["K"], num_heads=c["heads"])
    ref = OrderedDict((k, tuple(v.shape)) for k, v in ic.state_dict().items())
    assert list(ref.items()) == list(mini_head_shapes().items())


def test_icl_head_golden():
    g = golden("icl_head_mini")
    c = MINI
    state = synth.synth_state_dict(mini_head_shapes(), 11)
    P = R.make_params(OrderedDict(("h." + k, v) for k, v in state.items()))
    feats = [synth.synth_volume((c["B"], ch) + (r,) * 3, 20 + i).requires_grad_(True)
             for i, (ch, r) in enumerate(zip(c["in_chans"], c["res"]))]
    fm_l, q_l = R.inherent_consistent(P, "h", feats, None, "labeled", heads=c["heads"])
    fm_u, _ = R.inherent_consistent(P, "h", feats, [q.detach() for q in q_l], "unlabeled", heads=c["heads"])
    for i in range(3):
        assert_close(fm_l[i].detach(), g["fm_l%d" % i], TOL, "fm_l%d" % i)
        assert_close(fm_u[i].detach(), g["fm_u%d" % i], TOL, "fm_u%d" % i)
        assert_close(q_l[i].detach(), g["q_l%d" % i], TOL, "q_l%d" % i)
    loss = sum((f ** 2).mean() for f in fm_l) + sum((f ** 2).mean() for f in fm_u) + sum((q ** 2).mean() for q in q_l)
    assert abs(loss.item() - float(g["loss"])) < 1e-4 * float(g["loss"])
    loss.backward()
    none = set(str(s) for s in g["grad_none"])
    for k, p in P.items():
        name = k[2:]
        if not p.requires_grad:
            continue
        if name in none:
            assert p.grad is None, name
        elif "g/" + name in g.files:
            assert_close(p.grad, g["g/" + name], 2e-4, name, abs_floor=1e-6)
        else:
            check_summary(p.grad, g["gsum/" + name], g["gval/" + name], 2e-4, name)
    for i in range(3):
        assert_close(feats[i].grad, g["dfeat%d" % i], 2e-4, "dfeat%d" % i)
    for k in g.files:
        if k.startswith("stat/") and "running" in k:
            assert_close(P["h." + k[5:]], g[k], 1e-5, k)


@pytest.mark.parametrize("name", ["losses_k2", "losses_k5"])
def test_losses_golden(name):
    g = golden(name)
    K = int(g["K"])
    B = 2
    labels = synth.synth_blobs((B, 96, 96, 96), K, 31)
    final_lab = synth.synth_volume((B, K, 96, 96, 96), 32).requires_grad_(True)
    final_unlab = synth.synth_volume((B, K, 96, 96, 96), 33)
    mk = lambda s: [synth.synth_volume((B, K, r, r, r), s + i).mul_(2.0) for i, r in enumerate((6, 12, 24))]
    fms = [t.requires_grad_(True) for t in mk(40)]
    fms2 = [t.requires_grad_(True) for t in mk(50)]
    fms3 = mk(60)
    L = R.icl_losses((final_lab, final_unlab, fms, fms2, fms3), labels, K)
    for k in ("ce", "dice", "aux", "pse", "cons", "total"):
        assert abs(L[k].item() - float(g[k])) <= 2e-6 * max(1.0, abs(float(g[k]))), k
    L["total"].backward()
    check_summary(final_lab.grad, g["dfinal_sum"], g["dfinal_val"], 1e-4, "dfinal", n=256, abs_floor=1e-9)
    for i in range(3):
        assert_close(fms[i].grad, g["daux%d" % i], 1e-4, "daux%d" % i)
        assert_close(fms2[i].grad, g["dpse%d" % i], 1e-4, "dpse%d" % i)


def test_sliding_window_golden():
    g = golden("sliding_window")
    P = R.make_params(synth.synth_state_dict(backbone_shapes(4, 2, 1), 77), requires_grad=False)
    image = synth.synth_volume((120, 104, 90), 78).numpy()
    label = R.test_single_case(lambda p: R.unet_3d_forward(P, p), image, 64, 64, (96, 96, 96), 2)
    want = np.unpackbits(g["label_bits"])[: label.size].reshape(label.shape)
    agree = float((label == want).mean())
    assert agree >= 0.9999, agree
    gt = synth.synth_blobs((120, 104, 90), 2, 79).numpy()
    dice, counts = R.dice_metric(want, gt)
    assert list(counts) == [int(v) for v in g["counts"]]
    assert dice == float(g["dice"])


@pytest.mark.parametrize("name", ["step_cfg2", "step_cfg3"])
def test_train_step_port_golden(name):
    """R.train_step_3d (the step bench.py's CPU arm times when no reference tree is reachable, kind "port") against the fixture the
    UNMODIFIED reference produced for the same full-size step: the five losses, per-voxel label maps, the grad-is-None set and every
    parameter gradient (785 M parameters, summarised)."""
    import json
    import os
    from collections import OrderedDict
    from helpers import GOLDEN, assert_argmax_agrees, check_summary
    g = golden(name)
    K = int(g["K"])
    keys = json.load(open(os.path.join(GOLDEN, "state_keys.json")))["unet_3D_icl_k2"]
    shapes = OrderedDict()
    for k, s in keys:
        s = list(s)
        if k in ("final.weight", "final.bias"):
            s[0] = K
        elif k.endswith("guided_Q"):
            s[1] = K
        shapes[k] = tuple(s)
    P = R.make_params(synth.synth_state_dict(shapes, 1337))
    x = synth.synth_volume((4, 1, 96, 96, 96), 1338)
    y = synth.synth_labels((4, 96, 96, 96), K, 1339)
    L, grads, outputs = R.train_step_3d(P, x, y, 2, K, weights=dict(zip(("dice", "ce", "aux", "pse", "cons"), (float(v) for v in g["weights"]))),
                                       rand=R.NoRand())
    for k in ("ce", "dice", "aux", "pse", "cons", "total"):
        assert abs(float(L[k]) - float(g[k])) <= 1e-5 * max(1.0, abs(float(g[k]))), (k, float(L[k]), float(g[k]))
    assert_argmax_agrees(outputs[0], g["final_lab_argmax_bits"], "final_lab", 0.9999)
    assert_argmax_agrees(outputs[1], g["final_unlab_argmax_bits"], "final_unlab", 0.9999)
    none = set(str(s) for s in g["grad_none"])
    assert set(k for k, v in grads.items() if v is None) == none
    for k, v in grads.items():
        if v is None:
            continue
        if k.endswith(".0.bias") and "sspa" not in k and "uscl" not in k:
            assert v.norm().item() < 1e-3, k   # conv bias in front of InstanceNorm: mathematically zero gradient
            continue
        check_summary(v, g["gsum/" + k], g["gval/" + k], 2e-4, k, abs_floor=1e-6)
