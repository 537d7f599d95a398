"""CPU tests of the Swin path (BASELINE config 4): the oracle restatement (oracle/restate_swin.py) against the
live reference and against the committed fixtures; the host logic of icl_b200's SwinUnet mirror with the kernels replaced by
torch stand-ins (tests/cpu_standins.py); state_dict interchange."""
import json
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

import cpu_standins
from helpers import GOLDEN, assert_close, check_summary, golden
from oracle import restate as R
from oracle import restate2d as R2
from oracle import restate_swin as RS
from oracle import synth
from oracle.make_golden import eval_dropout_only, swin_config


def _losses(o, y, n_lab, K=4):
    L = R2.icl_losses_2d(o, y[:n_lab], K)
    return L


def _check_step(gd, o, L, named_grads, tol_fwd, tol_grad):
    for nm in ("ce", "dice", "aux", "pse", "cons", "total"):
        assert abs(float(L[nm]) - float(gd[nm])) <= 1e-4 * max(abs(float(gd[nm])), 1e-3), (nm, float(L[nm]), float(gd[nm]))
    for nm, t in (("out_lab", o[0]), ("out_unlab", o[1])):
        check_summary(t, gd[nm + "_sum"], gd[nm + "_val"], tol_fwd, nm, n=4096)
    for j, nm in ((2, "maps_lab"), (3, "maps_unlab"), (4, "maps_consis")):
        for i in range(3):
            check_summary(o[j][i], gd["%s%d_sum" % (nm, i)], gd["%s%d_val" % (nm, i)], tol_fwd, "%s%d" % (nm, i), n=4096)
    none = set(str(s) for s in gd["grad_none"])
    for k, g in named_grads:
        if k in none:
            assert g is None, k
        else:
            assert g is not None, k
            check_summary(g, gd["gsum/" + k], gd["gval/" + k], tol_grad, k, abs_floor=2e-7)


def test_oracle_swin_step_vs_fixture():
    """oracle/restate_swin.py reproduces the reference fixture: logits, ICL maps, the five losses, every parameter gradient and the
    set of parameters the reference leaves without a gradient."""
    gd = golden("step_swin_b2")
    n_lab, n_unlab, seed = int(gd["n_lab"]), int(gd["n_unlab"]), int(gd["seed"])
    shapes = OrderedDict((k, tuple(s)) for k, s in json.load(open(os.path.join(GOLDEN, "state_keys.json")))["swin_unet_k4"])
    state = synth.synth_state_dict(shapes, seed)
    P = R.make_params(state)
    n = n_lab + n_unlab
    x = synth.synth_volume((n, 1, 224, 224), seed + 1)
    y = synth.synth_labels((n, 224, 224), 4, seed + 2)
    o = RS.swin_unet_forward(P, x[:n_lab], x[n_lab:])
    L = _losses(o, y, n_lab)
    L["total"].backward()
    params = json.load(open(os.path.join(GOLDEN, "state_keys.json")))["swin_unet_k4_params"]
    _check_step(gd, o, L, [(k, P[k].grad) for k in params], 2e-5, 5e-4)


@pytest.mark.reference
def test_oracle_swin_vs_live_reference():
    """Forward and gradients of the whole SwinUnet ICL model against the unmodified reference, DropPath replayed in train mode."""
    from oracle import ref_import
    ns = ref_import.load()
    m = ns.SwinUnet(swin_config(), img_size=224, num_classes=4)
    synth.load_synth(m, 31)
    m.train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.eval()
    x = synth.synth_volume((4, 1, 224, 224), 32)
    torch.manual_seed(7)
    ref = m(x[:2], x[2:])
    loss = sum((t ** 2).mean() for t in ref[:2]) + sum((t ** 2).mean() for lst in ref[2:] for t in lst)
    loss.backward()
    P = R.make_params(OrderedDict((k, v.detach().clone()) for k, v in m.state_dict().items()))
    torch.manual_seed(7)
    out = RS.swin_unet_forward(P, x[:2], x[2:], rand=R.TorchRand())
    loss2 = sum((t ** 2).mean() for t in out[:2]) + sum((t ** 2).mean() for lst in out[2:] for t in lst)
    loss2.backward()
    assert_close(out[0].detach(), ref[0].detach(), 2e-5, "out_lab")
    assert_close(out[1].detach(), ref[1].detach(), 2e-5, "out_unlab")
    for j in (2, 3, 4):
        for a, b in zip(out[j], ref[j]):
            assert_close(a.detach(), b.detach(), 5e-5, "maps")
    for k, p in m.named_parameters():
        if p.grad is None:
            assert P[k].grad is None, k
        else:
            assert_close(P[k].grad, p.grad, 5e-4, k, abs_floor=1e-7)


@pytest.mark.reference
def test_swin_geometry_buffers_match_reference():
    """relative_position_index and the shift masks derived from coordinates equal the reference's buffers."""
    from oracle import ref_import
    ns = ref_import.load()
    for res, shift in (((14, 14), 3), ((56, 56), 3), ((28, 28), 3)):
        blk = ns.SwinTransformerBlock(96, res, 3, window_size=7, shift_size=shift)
        assert torch.equal(blk.attn.relative_position_index, RS.relative_position_index(7))
        assert torch.equal(blk.attn_mask, RS.shift_mask(res[0], res[1], 7, shift))


def test_swin_state_dict_keys_match_reference_fixture():
    from icl_b200.networks.vision_transformer import SwinUnet, swin_tiny_lite_config
    keys = json.load(open(os.path.join(GOLDEN, "state_keys.json")))
    m = SwinUnet(swin_tiny_lite_config(), img_size=224, num_classes=4)
    assert [[k, list(v.shape)] for k, v in m.state_dict().items()] == keys["swin_unet_k4"]
    assert [k for k, _ in m.named_parameters()] == keys["swin_unet_k4_params"]
    from icl_b200.networks.unet_icl import UNet_icl
    from icl_b200.networks.unet import UNet
    m2 = UNet_icl(1, 4)
    assert [[k, list(v.shape)] for k, v in m2.state_dict().items()] == keys["unet_icl_k4"]
    assert [k for k, _ in m2.named_parameters()] == keys["unet_icl_k4_params"]
    m3 = UNet(1, 4)
    assert [[k, list(v.shape)] for k, v in m3.state_dict().items()] == keys["unet2d_k4"]
    # geometry buffers are functions of the window layout, identical to the oracle's (and, by the test above, the reference's)
    for k, v in m.state_dict().items():
        if k.endswith("relative_position_index"):
            assert torch.equal(v, RS.relative_position_index(7))
    assert torch.equal(m.swin_unet.layers[0].blocks[1].attn_mask, RS.shift_mask(56, 56, 7, 3))


def test_swin_host_logic_with_standins(monkeypatch):
    """icl_b200's SwinUnet module tree driven on CPU with torch stand-ins for the kernels: same logits, ICL maps, losses-ready
    outputs, gradients and None-gradient set as the reference fixture.  (The kernels themselves are tested on the GPU.)"""
    cpu_standins.install(monkeypatch)
    from icl_b200.networks.vision_transformer import SwinUnet, swin_tiny_lite_config
    gd = golden("step_swin_b2")
    n_lab, n_unlab, seed = int(gd["n_lab"]), int(gd["n_unlab"]), int(gd["seed"])
    net = SwinUnet(swin_tiny_lite_config(), img_size=224, num_classes=4)
    synth.load_synth(net, seed)
    net.train()
    eval_dropout_only(net)
    n = n_lab + n_unlab
    x = synth.synth_volume((n, 1, 224, 224), seed + 1)
    y = synth.synth_labels((n, 224, 224), 4, seed + 2)
    o = net(x[:n_lab], x[n_lab:])
    L = _losses(o, y, n_lab)
    L["total"].backward()
    _check_step(gd, o, L, [(k, p.grad) for k, p in net.named_parameters()], 2e-5, 5e-4)
    # inference path: logits only, 1-channel input repeated to 3 (vision_transformer.py:91-95)
    with torch.no_grad():
        out = net(x[:n_lab], inference=True)
    check_summary(out, gd["out_lab_sum"], gd["out_lab_val"], 2e-5, "inference logits", n=4096)
