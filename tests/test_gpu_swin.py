"""GPU parity tests of the Swin path (BASELINE config 4): the window-attention kernel against the oracle's
torch formulation, and the SwinUnet ICL forward + five losses + backward against fixtures produced by the unmodified reference."""
import pytest
import torch

from helpers import assert_argmax_agrees, assert_close, check_summary, golden
from oracle import restate_swin as RS
from oracle import synth
from oracle.make_golden import eval_dropout_only

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,H,W,nH,ws,shift", [(2, 14, 14, 3, 7, 0), (2, 14, 14, 3, 7, 3), (1, 56, 56, 3, 7, 3), (3, 7, 7, 24, 7, 0),
                                               (2, 28, 14, 6, 7, 3), (2, 8, 16, 2, 4, 2), (1, 16, 16, 1, 8, 4)])
def test_window_attention_kernel_vs_torch(B, H, W, nH, ws, shift):
    """Forward, dqkv and the relative-position-bias gradient of the fused (shifted-)window attention kernel."""
    import icl_b200.functional as Fn
    C = nH * 32
    g = torch.Generator(device="cuda").manual_seed(H * 100 + W + shift)
    qkv = torch.randn(B, H * W, 3 * C, device="cuda", generator=g)
    table = torch.randn((2 * ws - 1) ** 2, nH, device="cuda", generator=g) * 0.5
    dout = torch.randn(B, H * W, C, device="cuda", generator=g)
    q1, t1 = qkv.clone().requires_grad_(True), table.clone().requires_grad_(True)
    out = Fn.window_attention(q1, t1, H, W, nH, ws, shift)
    out.backward(dout)
    q2, t2 = qkv.double().requires_grad_(True), table.double().requires_grad_(True)
    ref = RS.window_attention_tokens(q2, t2, H, W, nH, ws, shift)
    ref.backward(dout.double())
    assert_close(out.detach().cpu(), ref.detach().cpu(), 1e-5, "window attention fwd")
    assert_close(q1.grad.cpu(), q2.grad.cpu(), 1e-5, "dqkv")
    assert_close(t1.grad.cpu(), t2.grad.cpu(), 2e-5, "dtable")


def test_window_attention_rejects_bad_geometry():
    import icl_b200.functional as Fn
    qkv = torch.zeros(1, 14 * 14, 3 * 64, device="cuda")
    with pytest.raises(RuntimeError):
        Fn.window_attention(qkv, torch.zeros(169, 1, device="cuda"), 14, 14, 1, 7, 0)   # head_dim 64
    with pytest.raises(RuntimeError):
        Fn.window_attention(torch.zeros(1, 15 * 15, 96, device="cuda"), torch.zeros(169, 1, device="cuda"), 15, 15, 1, 7, 0)


def _swin_step(name, tol_fwd, tol_grad):
    from icl_b200.networks.vision_transformer import SwinUnet, swin_tiny_lite_config
    from icl_b200.utils import losses as L
    gd = golden(name)
    K, n_lab, n_unlab, seed = int(gd["K"]), int(gd["n_lab"]), int(gd["n_unlab"]), int(gd["seed"])
    net = SwinUnet(swin_tiny_lite_config(), img_size=224, num_classes=K)
    synth.load_synth(net, seed)
    net.cuda().train()
    eval_dropout_only(net)
    n = n_lab + n_unlab
    x = synth.synth_volume((n, 1, 224, 224), seed + 1).cuda()
    y = synth.synth_labels((n, 224, 224), K, seed + 2).cuda()
    o = net(x[:n_lab], x[n_lab:])
    ce = L.CrossEntropyLoss()(o[0], y[:n_lab].long())
    dice = L.DiceLoss(K)(o[0], y[:n_lab].unsqueeze(1), softmax=True)
    aux = L.AuxLoss(K)(o[2], y[:n_lab])
    pse = L.PseudoSoftLoss(K)(o[3], o[1])
    cons = L.softmax_mse_loss(o[3], o[4])
    total = ce + dice + aux + pse + 50 * cons
    for nm, v in (("ce", ce), ("dice", dice), ("aux", aux), ("pse", pse), ("cons", cons), ("total", total)):
        assert abs(v.item() - float(gd[nm])) <= 1e-3 * max(abs(float(gd[nm])), 1e-3), (nm, v.item(), float(gd[nm]))
    import numpy as np
    for nm, t in (("out_lab", o[0]), ("out_unlab", o[1])):
        check_summary(t, gd[nm + "_sum"], gd[nm + "_val"], tol_fwd, nm, n=4096)
        assert_argmax_agrees(t, gd[nm + "_argmax_bits"], nm)  # per voxel, >= 99.9 %
    for j, nm in ((2, "maps_lab"), (3, "maps_unlab"), (4, "maps_consis")):
        for i in range(3):
            check_summary(o[j][i], gd["%s%d_sum" % (nm, i)], gd["%s%d_val" % (nm, i)], tol_fwd, "%s%d" % (nm, i), n=4096)
    total.backward()
    none = set(str(s) for s in gd["grad_none"])
    for k, p in net.named_parameters():
        if k in none:
            assert p.grad is None, k
        else:
            assert p.grad is not None, k
            check_summary(p.grad, gd["gsum/" + k], gd["gval/" + k], tol_grad, k, abs_floor=2e-6)


def test_swin_unet_icl_step_small_golden():
    """1 labeled + 1 unlabeled 1x224x224 slice.  The backbone is smooth (GELU / softmax), so its gradients keep a tight bound;
    the ICL heads contain BatchNorm -> ReLU, whose derivative flips bound the rest (see tests/test_gpu_2d.py)."""
    _swin_step("step_swin_b2", 1e-3, 3e-2)


def test_swin_unet_icl_step_cfg4_golden():
    """BASELINE config 4: SwinUnet(swin-tiny lite, 224, 4 classes), 8 labeled + 8 unlabeled slices."""
    _swin_step("step_cfg4", 1e-3, 3e-2)


def test_swin_inference_path():
    from icl_b200.networks.vision_transformer import SwinUnet, swin_tiny_lite_config
    gd = golden("step_swin_b2")
    seed = int(gd["seed"])
    net = SwinUnet(swin_tiny_lite_config(), img_size=224, num_classes=4)
    synth.load_synth(net, seed)
    net.cuda().eval()
    x = synth.synth_volume((2, 1, 224, 224), seed + 1).cuda()
    with torch.no_grad():
        out = net(x[:1], inference=True)
    assert tuple(out.shape) == (1, 4, 224, 224)
    check_summary(out, gd["out_lab_sum"], gd["out_lab_val"], 1e-3, "inference logits", n=4096)
