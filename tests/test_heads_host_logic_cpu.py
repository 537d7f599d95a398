"""CPU: the HOST LOGIC of the ICL-head mirrors (icl_b200/networks/unet_3D_icl.py and unet_icl.py: module wiring, raw reshapes,
DropPath ordering, which parameters receive gradients) with the kernels replaced by torch stand-ins (tests/cpu_standins.py),
against the fixtures the unmodified reference produced.  The kernels themselves are tested on the GPU against the same fixtures."""
import torch

import cpu_standins
from helpers import assert_close, check_summary, golden
from oracle import synth
from oracle.make_golden import MINI, MINI2D, eval_dropout_only


def _run(ic, feats, g, tol):
    fm_l, q_l = ic(feats, None, "labeled")
    fm_u, q_u = ic(feats, [q.detach() for q in q_l], "unlabeled")
    for i in range(3):
        assert_close(fm_l[i].detach(), g["fm_l%d" % i], tol, "fm_l%d" % i)
        assert_close(fm_u[i].detach(), g["fm_u%d" % i], tol, "fm_u%d" % i)
        assert_close(q_l[i].detach(), g["q_l%d" % i], tol, "q_l%d" % i)
    loss = sum((f ** 2).mean() for f in fm_l) + sum((f ** 2).mean() for f in fm_u) + sum((q ** 2).mean() for q in q_l)
    assert abs(loss.item() - float(g["loss"])) < 1e-4 * float(g["loss"])
    loss.backward()
    none = set(str(s) for s in g["grad_none"])
    for k, p in ic.named_parameters():
        if k in none:
            assert p.grad is None, k
        elif "g/" + k in g.files:
            assert_close(p.grad, g["g/" + k], 5e-4, k, abs_floor=1e-6)
        else:
            check_summary(p.grad, g["gsum/" + k], g["gval/" + k], 5e-4, k)
    for i in range(3):
        assert_close(feats[i].grad, g["dfeat%d" % i], 5e-4, "dfeat%d" % i)


def test_icl_heads_3d_host_logic(monkeypatch):
    cpu_standins.install(monkeypatch)
    from icl_b200.networks.unet_3D_icl import InherentConsistent
    c = MINI
    ic = InherentConsistent(in_chans=c["in_chans"], depths=(2, 2, 2), patch_size=(2, 2, 2), input_resolution=c["res"],
                            num_classes=c["K"], num_heads=c["heads"])
    synth.load_synth(ic, 11)
    ic.train()
    eval_dropout_only(ic)
    feats = [synth.synth_volume((c["B"], ch) + (r,) * 3, 20 + i).requires_grad_(True)
             for i, (ch, r) in enumerate(zip(c["in_chans"], c["res"]))]
    _run(ic, feats, golden("icl_head_mini"), 5e-5)


def test_icl_heads_2d_host_logic(monkeypatch):
    cpu_standins.install(monkeypatch)
    from icl_b200.networks.unet_icl import InherentConsistent
    c = MINI2D
    ic = InherentConsistent(in_chans=c["in_chans"], depths=(2, 2, 2), patch_size=(2, 2), input_resolution=c["res"], num_classes=c["K"],
                            num_heads=c["heads"])
    synth.load_synth(ic, 31)
    ic.train()
    eval_dropout_only(ic)
    feats = [synth.synth_volume((c["B"], ch, r, r), 40 + i).requires_grad_(True) for i, (ch, r) in enumerate(zip(c["in_chans"], c["res"]))]
    _run(ic, feats, golden("icl_head2d_mini"), 5e-5)


def test_need_queries_false_prunes_the_dead_branch(monkeypatch):
    """unlabeled mode with need_queries=False (what UNet_icl / unet_3D_icl / SwinUnet pass for the uscl call): identical maps, and the
    proxy-update parameters stay without gradient exactly as in the reference's graph."""
    cpu_standins.install(monkeypatch)
    from icl_b200.networks.unet_icl import InherentConsistent
    c = MINI2D
    ic = InherentConsistent(in_chans=c["in_chans"], depths=(2, 2, 2), patch_size=(2, 2), input_resolution=c["res"], num_classes=c["K"],
                            num_heads=c["heads"])
    synth.load_synth(ic, 31)
    ic.train()
    eval_dropout_only(ic)
    feats = [synth.synth_volume((c["B"], ch, r, r), 40 + i) for i, (ch, r) in enumerate(zip(c["in_chans"], c["res"]))]
    with torch.no_grad():
        _, q_l = ic(feats, None, "labeled")
    a, qa = ic(feats, q_l, "unlabeled")
    b, qb = ic(feats, q_l, "unlabeled", need_queries=False)
    assert len(qa) == 3 and len(qb) == 0
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    sum((m ** 2).mean() for m in b).backward()
    got_none = sorted(k for k, p in ic.named_parameters() if p.grad is None)
    assert any("attn.proj" in k for k in got_none) and any("query_convs" in k for k in got_none) and "guided_Q" in got_none
    assert not any("fc_kv" in k or "mlp2" in k or "attn_convs" in k for k in got_none)
