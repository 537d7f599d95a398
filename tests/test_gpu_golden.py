"""GPU parity tests against the fixtures produced by the UNMODIFIED reference (tests/golden, generator
oracle/make_golden.py): backbone logits + gradients, ICL heads, the five losses, one full config-2 training
step (785 M parameters) and the sliding-window label map.  Tolerances follow BASELINE.json north_star:
logits / losses within 1e-3 relative (we assert 5e-4 or tighter in "parity" mode), label maps >= 99.9 %."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from helpers import assert_argmax_agrees, assert_close, check_summary, golden
from oracle import synth
from oracle.make_golden import MINI

pytestmark = pytest.mark.gpu


def eval_dropout_only(model):
    for m in model.modules():
        if m.__class__.__name__ in ("Dropout", "DropPath"):
            m.eval()


@pytest.mark.parametrize("name", ["unet3d_fs4_k2_c1_32", "unet3d_fs16_k3_c2_32"])
def test_backbone_golden(name):
    from icl_b200.networks.unet_3D import unet_3D
    from icl_b200.utils import losses as L
    g = golden(name)
    fs, K, cin, size, seed, B = [int(v) for v in g["meta"]]
    net = unet_3D(feature_scale=fs, n_classes=K, in_channels=cin)
    synth.load_synth(net, seed)
    net.cuda().train()
    eval_dropout_only(net)
    x = synth.synth_volume((B, cin) + (size,) * 3, seed + 1).cuda()
    y = synth.synth_labels((B,) + (size,) * 3, K, seed + 2).cuda()
    logits = net(x)
    assert tuple(logits.shape) == (B, K, size, size, size)
    assert_close(logits.detach().cpu(), g["logits"], 5e-4, "logits")
    loss = L.CrossEntropyLoss()(logits, y) + L.DiceLoss(K)(torch.softmax(logits, 1), y.unsqueeze(1))
    assert abs(loss.item() - float(g["loss"])) < 5e-4 * float(g["loss"])
    loss.backward()
    for k, p in net.named_parameters():
        # conv biases in front of InstanceNorm have mathematically-zero gradients: absolute floor
        check_summary(p.grad, g["gsum/" + k], g["gval/" + k], 2e-3, k, abs_floor=2e-6)


def test_icl_head_golden():
    from icl_b200.networks.unet_3D_icl import InherentConsistent
    g = golden("icl_head_mini")
    c = MINI
    ic = InherentConsistent(in_chans=c["in_chans"], depths=(2, 2, 2), patch_size=(2, 2, 2), input_resolution=c["res"],
                            num_classes=c["K"], num_heads=c["heads"])
    synth.load_synth(ic, 11)
    ic.cuda().train()
    eval_dropout_only(ic)
    feats = [synth.synth_volume((c["B"], ch) + (r,) * 3, 20 + i).cuda().requires_grad_(True)
             for i, (ch, r) in enumerate(zip(c["in_chans"], c["res"]))]
    fm_l, q_l = ic(feats, None, "labeled")
    fm_u, q_u = ic(feats, [q.detach() for q in q_l], "unlabeled")
    for i in range(3):
        assert_close(fm_l[i].detach().cpu(), g["fm_l%d" % i], 1e-4, "fm_l%d" % i)
        assert_close(fm_u[i].detach().cpu(), g["fm_u%d" % i], 1e-4, "fm_u%d" % i)
        assert_close(q_l[i].detach().cpu(), g["q_l%d" % i], 1e-4, "q_l%d" % i)
    loss = sum((f ** 2).mean() for f in fm_l) + sum((f ** 2).mean() for f in fm_u) + sum((q ** 2).mean() for q in q_l)
    assert abs(loss.item() - float(g["loss"])) < 2e-4 * float(g["loss"])
    loss.backward()
    none = set(str(s) for s in g["grad_none"])
    for k, p in ic.named_parameters():
        if k in none:
            assert p.grad is None, k
        elif "g/" + k in g.files:
            assert_close(p.grad.cpu(), g["g/" + k], 5e-4, k, abs_floor=1e-6)
        else:
            check_summary(p.grad, g["gsum/" + k], g["gval/" + k], 5e-4, k)
    for i in range(3):
        assert_close(feats[i].grad.cpu(), g["dfeat%d" % i], 5e-4, "dfeat%d" % i)
    sd = ic.state_dict()
    for k in g.files:
        if k.startswith("stat/"):
            assert_close(sd[k[5:]].cpu(), g[k], 1e-5, k)


@pytest.mark.parametrize("name", ["losses_k2", "losses_k5"])
def test_losses_golden(name):
    from icl_b200.utils import losses as L
    g = golden(name)
    K, B = int(g["K"]), 2
    labels = synth.synth_blobs((B, 96, 96, 96), K, 31).cuda()
    final_lab = synth.synth_volume((B, K, 96, 96, 96), 32).cuda().requires_grad_(True)
    final_unlab = synth.synth_volume((B, K, 96, 96, 96), 33).cuda()
    mk = lambda s: [synth.synth_volume((B, K, r, r, r), s + i).mul_(2.0).cuda() for i, r in enumerate((6, 12, 24))]
    fms = [t.requires_grad_(True) for t in mk(40)]
    fms2 = [t.requires_grad_(True) for t in mk(50)]
    fms3 = mk(60)
    ce = L.CrossEntropyLoss()(final_lab, labels)
    dice = L.DiceLoss(K)(torch.softmax(final_lab, 1), labels.unsqueeze(1))
    aux = L.AuxLoss3D(K)(fms, labels)
    pse = L.PseudoSoftLoss3D(K)(fms2, final_unlab)
    cons = L.softmax_mse_loss(fms2, fms3)
    total = dice + ce + aux + pse + 10 * cons
    for k, v in (("ce", ce), ("dice", dice), ("aux", aux), ("pse", pse), ("cons", cons), ("total", total)):
        assert abs(v.item() - float(g[k])) <= 1e-5 * max(1.0, abs(float(g[k]))), (k, v.item(), float(g[k]))
    total.backward()
    check_summary(final_lab.grad, g["dfinal_sum"], g["dfinal_val"], 2e-4, "dfinal", n=256, abs_floor=1e-9)
    for i in range(3):
        assert_close(fms[i].grad.cpu(), g["daux%d" % i], 2e-4, "daux%d" % i)
        assert_close(fms2[i].grad.cpu(), g["dpse%d" % i], 2e-4, "dpse%d" % i)


@pytest.mark.parametrize("name", ["step_cfg2", "step_cfg3"])
def test_full_step_golden(name):
    """BASELINE config 2: unet_3D_icl(n_classes=2, in_channels=1), 2 labeled + 2 unlabeled 96^3 patches, BraTS loss weights
    (train_inherent_consistent_unet_3D_BraTS.py:112); config 3: the same with n_classes=16 and the AMOS loss weights 1,1,1,0.1,10
    (train_inherent_consistent_unet_3D_AMOS22.py:230) — 128-row mlp2 GEMMs, K = 16 loss kernels."""
    from icl_b200.networks.unet_3D_icl import unet_3D_icl
    from icl_b200.utils import losses as L
    g = golden(name)
    K = int(g["K"])
    w = [float(v) for v in g["weights"]]
    net = unet_3D_icl(feature_scale=4, n_classes=K, in_channels=1)
    synth.load_synth(net, 1337)
    net.cuda().train()
    eval_dropout_only(net)
    x = synth.synth_volume((4, 1, 96, 96, 96), 1338).cuda()
    y = synth.synth_labels((4, 96, 96, 96), K, 1339).cuda()
    o = net(x[:2], x[2:])
    ce = L.CrossEntropyLoss()(o[0], y[:2])
    dice = L.DiceLoss(K)(torch.softmax(o[0], 1), y[:2].unsqueeze(1))
    aux = L.AuxLoss3D(K)(o[2], y[:2])
    pse = L.PseudoSoftLoss3D(K)(o[3], o[1])
    cons = L.softmax_mse_loss(o[3], o[4])
    total = w[0] * dice + w[1] * ce + w[2] * aux + w[3] * pse + w[4] * cons
    total.backward()
    torch.cuda.synchronize()
    for k, v in (("ce", ce), ("dice", dice), ("aux", aux), ("pse", pse), ("cons", cons), ("total", total)):
        assert abs(v.item() - float(g[k])) <= 1e-3 * abs(float(g[k])), (k, v.item(), float(g[k]))
    for nm, t in (("final_lab", o[0]), ("final_unlab", o[1])):
        check_summary(t, g[nm + "_sum"], g[nm + "_val"], 5e-4, nm, n=4096)
        assert_argmax_agrees(t, g[nm + "_argmax_bits"], nm)  # per voxel, >= 99.9 %
    for j, nm in ((2, "maps_lab"), (3, "maps_unlab"), (4, "maps_consis")):
        for i in range(3):
            t = o[j][i].detach()
            if "%s%d" % (nm, i) in g.files:
                assert_close(t.cpu(), g["%s%d" % (nm, i)], 1e-3, "%s%d" % (nm, i))
            else:
                check_summary(t, g["%s%d_sum" % (nm, i)], g["%s%d_val" % (nm, i)], 1e-3, "%s%d" % (nm, i), n=4096)
    none = set(str(s) for s in g["grad_none"])
    got_none = set(k for k, p in net.named_parameters() if p.grad is None)
    assert got_none == none, (sorted(got_none - none), sorted(none - got_none))
    bad = []
    for k, p in net.named_parameters():
        if p.grad is None:
            continue
        if k.endswith(".0.bias") and "sspa" not in k and "uscl" not in k:
            # conv bias in front of InstanceNorm: the gradient is mathematically zero; both sides hold only
            # rounding noise (reference l2 ~1e-5), so compare against an absolute bound instead of each other
            if p.grad.norm().item() > 1e-3 or float(g["gsum/" + k][1]) > 1e-3:
                bad.append("%s: zero-gradient bias has l2 %.3e (ref %.3e)" % (k, p.grad.norm().item(), float(g["gsum/" + k][1])))
            continue
        try:
            # sanity bound only (ReLU-derivative / MaxPool-winner flips between two correct implementations move individual gradient
            # entries by their full size; K = 16 with random labels has the noisier loss gradient): the backward
            # ARITHMETIC is pinned at 1e-3 by test_backbone_grads_mask_matched, where those discrete choices are held fixed
            check_summary(p.grad, g["gsum/" + k], g["gval/" + k], 5e-3 if K == 2 else 1e-2, k, abs_floor=5e-6)
        except AssertionError as e:
            bad.append(str(e))
    assert not bad, "\n".join(bad[:20])


def test_sliding_window_golden():
    from icl_b200 import inference
    from icl_b200.networks.unet_3D import unet_3D
    g = golden("sliding_window")
    net = unet_3D(feature_scale=4, n_classes=2, in_channels=1)
    synth.load_synth(net, 77)
    net.cuda().eval()
    image = synth.synth_volume((120, 104, 90), 78).numpy()
    label = inference.test_single_case(net, image, 64, 64, (96, 96, 96), num_classes=2)
    assert label.shape == (120, 104, 90) and label.dtype == np.int64
    want = np.unpackbits(g["label_bits"])[: label.size].reshape(label.shape)
    agree = float((label == want).mean())
    assert agree >= 0.999, agree
    gt = synth.synth_blobs((120, 104, 90), 2, 79)
    dice, counts = inference.dice_metric(torch.from_numpy(want.astype(np.int64)).cuda(), gt.cuda())
    assert list(counts) == [int(v) for v in g["counts"]]  # bit-exact on identical label maps
    assert dice == float(g["dice"])
