"""CPU, build container only: the oracle restatement (oracle/restate.py) against the LIVE, unmodified reference
imported from /root/reference/code (oracle/ref_import.py).  Skipped where the reference tree is absent (GPU box).
This is the pin of the oracle: the reference ships no tests or golden vectors of its own."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import ref_import, synth
from oracle import restate as R
from oracle.make_golden import MINI, eval_dropout_only
from helpers import assert_close

pytestmark = pytest.mark.reference
TOL = 2e-5


@pytest.fixture(scope="module")
def ns():
    return ref_import.load()


def _params_of(module):
    return R.make_params(OrderedDict((k, v.detach().clone()) for k, v in module.state_dict().items()))


@pytest.mark.parametrize("fs,K,cin,B", [(16, 3, 2, 1), (8, 2, 1, 2)])
def test_backbone_forward_backward(ns, fs, K, cin, B):
    """unet_3D.forward (networks/unet_3D.py:71-94) + CE + DiceLoss, logits and every parameter gradient."""
    m = ns.unet_3D(feature_scale=fs, n_classes=K, in_channels=cin)
    synth.load_synth(m, 5)
    m.train()
    eval_dropout_only(m)
    x = synth.synth_volume((B, cin, 32, 32, 32), 6)
    y = synth.synth_labels((B, 32, 32, 32), K, 7)
    ref = m(x)
    (torch.nn.CrossEntropyLoss()(ref, y) + ns.losses.DiceLoss(K)(torch.softmax(ref, 1), y.unsqueeze(1))).backward()
    P = _params_of(m)
    out = R.unet_3d_forward(P, x)
    (R.ce_loss(out, y) + R.dice_loss(torch.softmax(out, 1), y.unsqueeze(1), K)).backward()
    assert_close(out.detach(), ref.detach(), TOL, "logits")
    for k, p in m.named_parameters():
        assert_close(P[k].grad, p.grad, 2e-4, k, abs_floor=1e-6)


def test_backbone_dropout_masks_replayed(ns):
    """train mode: the oracle draws Dropout masks from the torch generator in the reference's order (unet_3D_icl.py:110,116)."""
    m = ns.unet_3D(feature_scale=16, n_classes=2, in_channels=1)
    synth.load_synth(m, 9)
    m.train()
    x = synth.synth_volume((1, 1, 32, 32, 32), 10)
    torch.manual_seed(123)
    ref = m(x)
    P = _params_of(m)
    torch.manual_seed(123)
    out = R.unet_3d_forward(P, x, rand=R.TorchRand())
    assert_close(out.detach(), ref.detach(), TOL, "logits with dropout")


def test_icl_heads_both_modes(ns):
    """InherentConsistent.forward (unet_3D_icl.py:202-242) labeled (SSPA) then unlabeled (USCL) with the updated proxies."""
    c = MINI
    ic = ns.InherentConsistent(in_chans=c["in_chans"], depths=(2, 2, 2), patch_size=(2, 2, 2), input_resolution=c["res"],
                               num_classes=c["K"], num_heads=c["heads"])
    synth.load_synth(ic, 11)
    ic.train()
    eval_dropout_only(ic)
    feats = [synth.synth_volume((c["B"], ch) + (r,) * 3, 20 + i) for i, (ch, r) in enumerate(zip(c["in_chans"], c["res"]))]
    P = R.make_params(OrderedDict(("h." + k, v.detach().clone()) for k, v in ic.state_dict().items()))
    fm_l, q_l = ic(feats, None, "labeled")
    fm_u, _ = ic(feats, [q.detach() for q in q_l], "unlabeled")
    o_l, oq_l = R.inherent_consistent(P, "h", feats, None, "labeled", heads=c["heads"])
    o_u, _ = R.inherent_consistent(P, "h", feats, [q.detach() for q in oq_l], "unlabeled", heads=c["heads"])
    for i in range(3):
        assert_close(o_l[i].detach(), fm_l[i].detach(), TOL, "sspa map %d" % i)
        assert_close(o_u[i].detach(), fm_u[i].detach(), TOL, "uscl map %d" % i)
        assert_close(oq_l[i].detach(), q_l[i].detach(), TOL, "proxy %d" % i)
    # BatchNorm running statistics are a side effect of the forward pass
    for k, v in ic.state_dict().items():
        if "running" in k:
            assert_close(P["h." + k], v, 1e-5, k)


@pytest.mark.parametrize("K", [2, 5])
def test_losses(ns, K):
    """CrossEntropy, DiceLoss (:195-231), AuxLoss3D (:254-271), PseudoSoftLoss3D (:287-299), softmax_mse_loss (:68-90)."""
    L = ns.losses
    B = 2
    labels = synth.synth_blobs((B, 96, 96, 96), K, 31)
    final_lab = synth.synth_volume((B, K, 96, 96, 96), 32)
    final_unlab = synth.synth_volume((B, K, 96, 96, 96), 33)
    mk = lambda s: [synth.synth_volume((B, K, r, r, r), s + i).mul_(2.0) for i, r in enumerate((6, 12, 24))]
    fms, fms2, fms3 = mk(40), mk(50), mk(60)
    want = dict(ce=torch.nn.CrossEntropyLoss()(final_lab, labels), dice=L.DiceLoss(K)(torch.softmax(final_lab, 1), labels.unsqueeze(1)),
                aux=L.AuxLoss3D(K)(fms, labels), pse=L.PseudoSoftLoss3D(K)(fms2, final_unlab), cons=L.softmax_mse_loss(fms2, fms3))
    got = R.icl_losses((final_lab, final_unlab, fms, fms2, fms3), labels, K)
    for k, v in want.items():
        assert abs(got[k].item() - v.item()) <= 2e-6 * max(1.0, abs(v.item())), k


def test_sliding_window_driver(ns):
    """test_single_case (test_3D_BraTS.py:79-142): padding, window grid, accumulation order, argmax."""
    tsc = ref_import.load_test_single_case()
    m = ns.unet_3D(feature_scale=16, n_classes=2, in_channels=1)
    synth.load_synth(m, 77)
    m.eval()
    image = synth.synth_volume((40, 50, 30), 78).numpy()
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        want = tsc(m, image, 16, 16, (32, 32, 32), num_classes=2)
    finally:
        torch.Tensor.cuda = orig
    P = _params_of(m)
    with torch.no_grad():
        got = R.test_single_case(lambda p: R.unet_3d_forward(P, p), image, 16, 16, (32, 32, 32), 2)
    assert got.shape == want.shape
    assert float((got == want).mean()) >= 0.9999


# ------------------------------------------------------------------------------------------ 2D path (config 1)
MINI2D = dict(in_chans=(64, 32, 16), res=[4, 8, 16], heads=(4, 2, 1), K=3, B=2)


def test_unet2d_forward_backward(ns):
    """UNet (networks/unet.py:305-322; same blocks as unet_icl.py): BatchNorm2d batch statistics, LeakyReLU, bilinear
    align_corners=True upsampling; logits, every parameter gradient and the running statistics."""
    from oracle import restate2d as R2
    m = ns.UNet(in_chns=1, class_num=4)
    synth.load_synth(m, 21)
    m.train()
    eval_dropout_only(m)
    x = synth.synth_volume((4, 1, 64, 64), 22)
    y = synth.synth_labels((4, 64, 64), 4, 23)
    P = _params_of(m)
    ref = m(x)
    (torch.nn.CrossEntropyLoss()(ref, y) + ns.losses.DiceLoss(4)(ref, y.unsqueeze(1), softmax=True)).backward()
    out = R2.unet2d_forward(P, x)
    (R.ce_loss(out, y) + R.dice_loss(out, y.unsqueeze(1), 4, softmax=True)).backward()
    assert_close(out.detach(), ref.detach(), TOL, "2D logits")
    for k, p in m.named_parameters():
        assert_close(P[k].grad, p.grad, 2e-4, k, abs_floor=1e-6)
    for k, v in m.state_dict().items():
        if "running" in k:
            assert_close(P[k], v, 1e-5, k)


def test_unet2d_dropout_masks_replayed(ns):
    """train mode: Dropout(p) inside every encoder ConvBlock draws from the torch generator in the reference's order."""
    from oracle import restate2d as R2
    m = ns.UNet(in_chns=1, class_num=2)
    synth.load_synth(m, 24)
    m.train()
    x = synth.synth_volume((2, 1, 32, 32), 25)
    P = _params_of(m)
    torch.manual_seed(7)
    ref = m(x)
    torch.manual_seed(7)
    out = R2.unet2d_forward(P, x, rand=R.TorchRand())
    assert_close(out.detach(), ref.detach(), TOL, "2D logits with dropout")


def test_icl_heads_2d(ns):
    """InherentConsistent with spatial_dims=2 (unet_icl.py:254-343), labeled then unlabeled mode."""
    from oracle import restate2d as R2
    c = MINI2D
    ic = ns.InherentConsistent2d(in_chans=c["in_chans"], depths=(2, 2, 2), patch_size=(2, 2), input_resolution=c["res"],
                                 num_classes=c["K"], num_heads=c["heads"])
    synth.load_synth(ic, 31)
    ic.train()
    eval_dropout_only(ic)
    feats = [synth.synth_volume((c["B"], ch, r, r), 40 + i) for i, (ch, r) in enumerate(zip(c["in_chans"], c["res"]))]
    P = R.make_params(OrderedDict(("h." + k, v.detach().clone()) for k, v in ic.state_dict().items()))
    fm_l, q_l = ic(feats, None, "labeled")
    fm_u, _ = ic(feats, [q.detach() for q in q_l], "unlabeled")
    o_l, oq_l = R2.inherent_consistent_2d(P, "h", feats, None, "labeled", heads=c["heads"])
    o_u, _ = R2.inherent_consistent_2d(P, "h", feats, [q.detach() for q in oq_l], "unlabeled", heads=c["heads"])
    for i in range(3):
        assert_close(o_l[i].detach(), fm_l[i].detach(), TOL, "2D sspa map %d" % i)
        assert_close(o_u[i].detach(), fm_u[i].detach(), TOL, "2D uscl map %d" % i)
        assert_close(oq_l[i].detach(), q_l[i].detach(), TOL, "2D proxy %d" % i)


def test_losses_2d(ns):
    """AuxLoss (:233-251), PseudoSoftLoss (:273-285) and the 2D use of DiceLoss(softmax=True) / softmax_mse_loss."""
    from oracle import restate2d as R2
    L = ns.losses
    B, K, S = 3, 4, 64
    labels = synth.synth_labels((B, S, S), K, 51)
    out_lab = synth.synth_volume((B, K, S, S), 52)
    out_unlab = synth.synth_volume((B, K, S, S), 53)
    mk = lambda s: [synth.synth_volume((B, K, r, r), s + i).mul_(2.0) for i, r in enumerate((8, 16, 32))]
    fms, fms2, fms3 = mk(60), mk(70), mk(80)
    want = dict(ce=torch.nn.CrossEntropyLoss()(out_lab, labels), dice=L.DiceLoss(K)(out_lab, labels.unsqueeze(1), softmax=True),
                aux=L.AuxLoss(K, resize=[S, S])(fms, labels), pse=L.PseudoSoftLoss(K, resize=[S, S])(fms2, out_unlab),
                cons=L.softmax_mse_loss(fms2, fms3))
    got = R2.icl_losses_2d((out_lab, out_unlab, fms, fms2, fms3), labels, K)
    for k, v in want.items():
        assert abs(got[k].item() - v.item()) <= 2e-6 * max(1.0, abs(v.item())), k
