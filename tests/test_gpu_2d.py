"""GPU parity tests of the 2D path (BASELINE config 1): kernels against torch CPU ops, modules and the
full UNet_icl step against fixtures produced by the unmodified reference (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import assert_close, check_summary, golden
from oracle import synth
from oracle.make_golden import MINI2D

pytestmark = pytest.mark.gpu


def g(seed):
    return torch.Generator().manual_seed(seed)


def eval_dropout_only(model):
    for m in model.modules():
        if m.__class__.__name__ in ("Dropout", "DropPath"):
            m.eval()


def test_conv_bn_leaky_block_vs_torch():
    """Conv2d -> BatchNorm2d(train) -> LeakyReLU with two concatenated sources: output, input / parameter gradients, running stats."""
    import icl_b200.functional2d as F2
    N, c0, c1, cout, H, W = 4, 16, 32, 32, 24, 16
    x0 = torch.randn(N, c0, H, W, generator=g(1), requires_grad=True)
    x1 = torch.randn(N, c1, H, W, generator=g(2), requires_grad=True)
    conv = torch.nn.Conv2d(c0 + c1, cout, 3, padding=1)
    bn = torch.nn.BatchNorm2d(cout)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(cout, generator=g(3)) + 0.5)
        bn.bias.copy_(torch.randn(cout, generator=g(4)) * 0.3)
    dy = torch.randn(N, cout, H, W, generator=g(5))
    ref = F.leaky_relu(bn(conv(torch.cat([x0, x1], 1))), 0.01)
    ref.backward(dy)
    import copy
    conv_c, bn_c = copy.deepcopy(conv).cuda(), torch.nn.BatchNorm2d(cout).cuda()
    with torch.no_grad():
        bn_c.weight.copy_(bn.weight); bn_c.bias.copy_(bn.bias)
    for p in list(conv_c.parameters()) + list(bn_c.parameters()):
        p.grad = None
    a0, a1 = x0.detach().cuda().requires_grad_(True), x1.detach().cuda().requires_grad_(True)
    out = F2.conv_bn_act(a0, a1, conv_c, bn_c, 0.01)
    out.backward(dy.cuda())
    assert_close(out.detach().cpu(), ref.detach(), 2e-4, "block fwd")
    assert_close(a0.grad.cpu(), x0.grad, 1e-3, "dx0")
    assert_close(a1.grad.cpu(), x1.grad, 1e-3, "dx1")
    assert_close(conv_c.weight.grad.cpu(), conv.weight.grad, 1e-3, "dw")
    assert_close(bn_c.weight.grad.cpu(), bn.weight.grad, 1e-3, "dgamma")
    assert_close(bn_c.bias.grad.cpu(), bn.bias.grad, 1e-3, "dbeta")
    assert_close(bn_c.running_mean.cpu(), bn.running_mean, 1e-4, "running_mean", abs_floor=1e-6)
    assert_close(bn_c.running_var.cpu(), bn.running_var, 1e-4, "running_var")
    assert int(bn_c.num_batches_tracked) == 1


def test_maxpool2d_and_bilinear_ac_vs_torch():
    import icl_b200.functional2d as F2
    x = torch.randint(0, 3, (3, 16, 8, 12), generator=g(6)).float().requires_grad_(True)  # many exact ties
    ref = F.max_pool2d(x, 2)
    dout = torch.randn(ref.shape, generator=g(7))
    ref.backward(dout)
    xc = x.detach().cuda().requires_grad_(True)
    out = F2.max_pool2d(xc)
    out.backward(dout.cuda())
    assert torch.equal(out.detach().cpu(), ref.detach()) and torch.equal(xc.grad.cpu(), x.grad)
    for C, h, w in ((8, 5, 7), (3, 4, 4), (16, 1, 6)):
        y = torch.randn(2, C, h, w, generator=g(8), requires_grad=True)
        up = F.interpolate(y, scale_factor=2, mode="bilinear", align_corners=True)
        du = torch.randn(up.shape, generator=g(9))
        up.backward(du)
        yc = y.detach().cuda().requires_grad_(True)
        o = F2.upsample2x_ac(yc)
        o.backward(du.cuda())
        assert_close(o.detach().cpu(), up.detach(), 1e-6, "bilinear ac fwd")
        assert_close(yc.grad.cpu(), y.grad, 1e-5, "bilinear ac bwd")


def test_unet2d_golden():
    from icl_b200.networks.unet import UNet
    from icl_b200.utils import losses as L
    gd = golden("unet2d_k4_64")
    K, size, seed, B = [int(v) for v in gd["meta"]]
    net = UNet(1, K)
    synth.load_synth(net, seed)
    net.cuda().train()
    eval_dropout_only(net)
    x = synth.synth_volume((B, 1, size, size), seed + 1).cuda()
    y = synth.synth_labels((B, size, size), K, seed + 2).cuda()
    logits = net(x)
    assert tuple(logits.shape) == (B, K, size, size)
    assert_close(logits.detach().cpu(), gd["logits"], 5e-4, "2D logits")
    loss = L.CrossEntropyLoss()(logits, y) + L.DiceLoss(K)(logits, y.unsqueeze(1), softmax=True)
    assert abs(loss.item() - float(gd["loss"])) < 5e-4 * float(gd["loss"])
    loss.backward()
    for k, p in net.named_parameters():
        check_summary(p.grad, gd["gsum/" + k], gd["gval/" + k], 3e-3, k, abs_floor=2e-6)
    for k, v in net.state_dict().items():
        if "running" in k:
            assert_close(v.cpu(), gd["stat/" + k], 1e-4, k, abs_floor=1e-6)


def test_icl_head2d_golden():
    from icl_b200.networks.unet_icl import InherentConsistent
    gd = golden("icl_head2d_mini")
    c = MINI2D
    ic = InherentConsistent(in_chans=c["in_chans"], depths=(2, 2, 2), patch_size=(2, 2), input_resolution=c["res"], num_classes=c["K"],
                            num_heads=c["heads"])
    synth.load_synth(ic, 31)
    ic.cuda().train()
    eval_dropout_only(ic)
    feats = [synth.synth_volume((c["B"], ch, r, r), 40 + i).cuda().requires_grad_(True) for i, (ch, r) in enumerate(zip(c["in_chans"], c["res"]))]
    fm_l, q_l = ic(feats, None, "labeled")
    fm_u, q_u = ic(feats, [q.detach() for q in q_l], "unlabeled")
    for i in range(3):
        assert_close(fm_l[i].detach().cpu(), gd["fm_l%d" % i], 2e-4, "fm_l%d" % i)
        assert_close(fm_u[i].detach().cpu(), gd["fm_u%d" % i], 2e-4, "fm_u%d" % i)
        assert_close(q_l[i].detach().cpu(), gd["q_l%d" % i], 2e-4, "q_l%d" % i)
    loss = sum((f ** 2).mean() for f in fm_l) + sum((f ** 2).mean() for f in fm_u) + sum((q ** 2).mean() for q in q_l)
    assert abs(loss.item() - float(gd["loss"])) < 2e-4 * float(gd["loss"])
    loss.backward()
    none = set(str(s) for s in gd["grad_none"])
    for k, p in ic.named_parameters():
        if k in none:
            assert p.grad is None, k
        else:
            assert_close(p.grad.cpu(), gd["g/" + k], 1e-3, k, abs_floor=1e-6)
    for i in range(3):
        assert_close(feats[i].grad.cpu(), gd["dfeat%d" % i], 1e-3, "dfeat%d" % i)


def test_losses2d_golden():
    from icl_b200.utils import losses as L
    gd = golden("losses2d_k4")
    K, S, B = [int(v) for v in gd["meta"]]
    labels = synth.synth_labels((B, S, S), K, 51).cuda()
    out_lab = synth.synth_volume((B, K, S, S), 52).cuda().requires_grad_(True)
    out_unlab = synth.synth_volume((B, K, S, S), 53).cuda()
    mk = lambda s: [synth.synth_volume((B, K, r, r), s + i).mul_(2.0).cuda() for i, r in enumerate((8, 16, 32))]
    fms = [t.requires_grad_(True) for t in mk(60)]
    fms2 = [t.requires_grad_(True) for t in mk(70)]
    fms3 = mk(80)
    ce = L.CrossEntropyLoss()(out_lab, labels)
    dice = L.DiceLoss(K)(out_lab, labels.unsqueeze(1), softmax=True)
    aux = L.AuxLoss(K, resize=[S, S])(fms, labels)
    pse = L.PseudoSoftLoss(K, resize=[S, S])(fms2, out_unlab)
    cons = L.softmax_mse_loss(fms2, fms3)
    for nm, v in (("ce", ce), ("dice", dice), ("aux", aux), ("pse", pse), ("cons", cons)):
        assert abs(v.item() - float(gd[nm])) <= 1e-5 * max(1.0, abs(float(gd[nm]))), nm
    (ce + dice + aux + pse + 50 * cons).backward()
    assert_close(out_lab.grad.cpu(), gd["dout"], 1e-4, "dout")
    for i in range(3):
        assert_close(fms[i].grad.cpu(), gd["daux%d" % i], 1e-4, "daux%d" % i)
        assert_close(fms2[i].grad.cpu(), gd["dpse%d" % i], 1e-4, "dpse%d" % i)


def test_full_step_2d_golden():
    """BASELINE config 1: UNet_icl(1, 4), 12 labeled + 12 unlabeled 1x256x256 slices, forward + 5 losses + backward."""
    from icl_b200.networks.unet_icl import UNet_icl
    from icl_b200.utils import losses as L
    gd = golden("step_cfg1")
    K = int(gd["K"])
    net = UNet_icl(1, K)
    synth.load_synth(net, 1337)
    net.cuda().train()
    eval_dropout_only(net)
    x = synth.synth_volume((24, 1, 256, 256), 1338).cuda()
    y = synth.synth_labels((24, 256, 256), K, 1339).cuda()
    o = net(x[:12], x[12:])
    ce = L.CrossEntropyLoss()(o[0], y[:12].long())
    dice = L.DiceLoss(K)(o[0], y[:12].unsqueeze(1), softmax=True)
    aux = L.AuxLoss(K, resize=[256, 256])(o[2], y[:12])
    pse = L.PseudoSoftLoss(K, resize=[256, 256])(o[3], o[1])
    cons = L.softmax_mse_loss(o[3], o[4])
    total = ce + dice + aux + pse + 50 * cons
    for nm, v in (("ce", ce), ("dice", dice), ("aux", aux), ("pse", pse), ("cons", cons), ("total", total)):
        assert abs(v.item() - float(gd[nm])) <= 1e-3 * max(abs(float(gd[nm])), 1e-3), (nm, v.item(), float(gd[nm]))
    for nm, t in (("out_lab", o[0]), ("out_unlab", o[1])):
        check_summary(t, gd[nm + "_sum"], gd[nm + "_val"], 1e-3, nm, n=4096)
        cnt = np.bincount(t.argmax(1).reshape(-1).cpu().numpy(), minlength=K)
        assert np.abs(cnt - gd[nm + "_argmax_count"]).sum() <= 2e-3 * t.numel() / K, nm  # >= 99.9 % of the label map agrees
    for j, nm in ((2, "maps_lab"), (3, "maps_unlab"), (4, "maps_consis")):
        for i in range(3):
            check_summary(o[j][i], gd["%s%d_sum" % (nm, i)], gd["%s%d_val" % (nm, i)], 2e-3, "%s%d" % (nm, i), n=4096)
    total.backward()
    none = set(str(s) for s in gd["grad_none"])
    for k, p in net.named_parameters():
        if k in none:
            assert p.grad is None, k
        else:
            assert p.grad is not None, k
            check_summary(p.grad, gd["gsum/" + k], gd["gval/" + k], 5e-3, k, abs_floor=2e-6)
