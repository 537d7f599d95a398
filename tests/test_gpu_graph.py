"""GPU: CUDA-graph replay of the whole training step (icl_b200/graph.py) against the eager step on the same state."""
import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu


def _eval_dropout_only(model):
    for m in model.modules():
        if m.__class__.__name__ in ("Dropout", "DropPath"):
            m.eval()


@pytest.mark.parametrize("fused", [False, True])
def test_graphed_step_matches_eager(fused):
    from icl_b200.graph import GraphedStep
    from icl_b200.networks.unet_3D_icl import unet_3D_icl
    from icl_b200.optim import SGD
    from icl_b200.utils import losses as L
    K = 2
    net = unet_3D_icl(feature_scale=4, n_classes=K, in_channels=1)
    synth.load_synth(net, 1337)
    net.cuda().train()
    _eval_dropout_only(net)  # identical arithmetic in both runs (masks would differ between the two RNG draws)
    opt = SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4, fused_factored=fused)
    aux, pse = L.AuxLoss3D(K), L.PseudoSoftLoss3D(K)
    x = synth.synth_volume((4, 1, 96, 96, 96), 1338).cuda()
    y = synth.synth_labels((4, 96, 96, 96), K, 1339).cuda()

    def step(xb, yb):
        o = net(xb[:2], xb[2:])
        ce, dice = L.seg_ce_dice(o[0], yb[:2])
        loss = dice + ce + aux(o[2], yb[:2]) + pse(o[3], o[1]) + 10 * L.softmax_mse_loss(o[3], o[4])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    watch = ["conv1.conv2.0.weight", "up_concat1.conv.conv1.0.weight", "final.weight", "sspa.class_decoders.2.mlp2.fc1.weight",
             "uscl.class_decoders.1.attn.fc_kv.weight"]
    params = dict(net.named_parameters())
    step(x, y)  # creates momentum buffers
    snap_p = {k: v.detach().clone() for k, v in net.state_dict().items()}
    snap_m = {id(p): opt.state[p]["momentum_buffer"].clone() for p in opt.state}

    def restore():
        with torch.no_grad():
            for k, v in net.state_dict().items():
                v.copy_(snap_p[k])
            for p in opt.state:
                opt.state[p]["momentum_buffer"].copy_(snap_m[id(p)])

    loss_e = step(x, y).item()
    after_e = {k: params[k].detach().clone() for k in watch}
    restore()
    g = GraphedStep(step, (x, y), opt, warmup=1)  # warm-up + capture both advance the state: restore before replaying
    restore()
    loss_g = g(x, y).item()
    assert abs(loss_g - loss_e) <= 1e-5 * abs(loss_e), (loss_g, loss_e)
    for k in watch:
        d = (params[k].detach() - after_e[k]).norm().item()
        ref = (after_e[k] - snap_p[k]).norm().item()
        # same update up to the order of fp32 atomics (split-K reductions are not bit-reproducible run to run)
        assert d <= 2e-2 * ref + 1e-6 * snap_p[k].norm().item(), (k, d, ref)
    # a second replay keeps training (state advances, loss changes)
    loss_g2 = g(x, y).item()
    assert loss_g2 != loss_g


def test_lanes_match_single_stream(monkeypatch):
    """The step with the ICL heads / fused updates / weight gradients on side streams (icl_b200/lanes.py, the default) against the
    same step issued on one stream (ICL_HEAD_LANES=0): same loss, same parameters after the update, same BatchNorm running
    statistics (the two sspa passes must update them in the reference's order)."""
    from icl_b200.graph import GraphedStep
    from icl_b200.networks.unet_3D_icl import unet_3D_icl
    from icl_b200.optim import SGD
    from icl_b200.utils import losses as L
    K = 2
    net = unet_3D_icl(feature_scale=4, n_classes=K, in_channels=1)
    synth.load_synth(net, 1337)
    net.cuda().train()
    _eval_dropout_only(net)
    opt = SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4, fused_factored=True)
    aux, pse = L.AuxLoss3D(K), L.PseudoSoftLoss3D(K)
    x = synth.synth_volume((4, 1, 96, 96, 96), 1338).cuda()
    y = synth.synth_labels((4, 96, 96, 96), K, 1339).cuda()

    def step(xb, yb):
        o = net(xb[:2], xb[2:])
        ce, dice = L.seg_ce_dice(o[0], yb[:2])
        loss = dice + ce + aux(o[2], yb[:2]) + pse(o[3], o[1]) + 10 * L.softmax_mse_loss(o[3], o[4])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    params = dict(net.named_parameters())
    step(x, y)  # creates momentum buffers
    snap_p = {k: v.detach().clone() for k, v in net.state_dict().items()}
    snap_m = {id(p): opt.state[p]["momentum_buffer"].clone() for p in opt.state}

    def restore():
        with torch.no_grad():
            for k, v in net.state_dict().items():
                v.copy_(snap_p[k])
            for p in opt.state:
                opt.state[p]["momentum_buffer"].copy_(snap_m[id(p)])

    def run(lanes_on, graphed):
        monkeypatch.setenv("ICL_HEAD_LANES", "1" if lanes_on else "0")
        restore()
        if graphed:
            g = GraphedStep(step, (x, y), opt, warmup=1)
            restore()
            loss = g(x, y).item()
            torch.cuda.synchronize()
            after = {k: v.detach().clone() for k, v in net.state_dict().items()}
            g.release()
        else:
            loss = step(x, y).item()
            after = {k: v.detach().clone() for k, v in net.state_dict().items()}
        return loss, after

    loss0, a0 = run(False, False)
    for graphed in (False, True):
        loss1, a1 = run(True, graphed)
        assert abs(loss1 - loss0) <= 1e-5 * abs(loss0), (graphed, loss1, loss0)
        for k in a0:
            if not a0[k].is_floating_point():
                assert torch.equal(a0[k], a1[k]), k
                continue
            d = (a1[k].float() - a0[k].float()).norm().item()
            upd = (a0[k].float() - snap_p[k].float()).norm().item()
            # same update up to the order of the fp32 / fp64 atomics (InstanceNorm statistics, split-K partials)
            assert d <= 2e-2 * upd + 1e-6 * snap_p[k].float().norm().item(), (graphed, k, d, upd)
