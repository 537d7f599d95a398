"""GPU, 2 ranks over real NCCL: the data-parallel training step (icl_b200.parallel: bucketed all-reduce from grad-ready hooks +
mlp2 factor all-gathers on the exchange stream + rank-R fused optimizer) against the parity definition of SURVEY.md §8e —
the R-rank parameters after a step equal a single-process step on the MEAN of the R rank-local gradients.
Skipped when fewer than 2 GPUs are visible (run with `gpurun --gpus 2`)."""
import os
import socket
import sys
import traceback

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _eval_dropout_only(model):
    for m in model.modules():
        if m.__class__.__name__ in ("Dropout", "DropPath"):
            m.eval()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _build(K, dev):
    from icl_b200.networks.unet_3D_icl import unet_3D_icl
    from icl_b200.utils import synth
    net = unet_3D_icl(feature_scale=4, n_classes=K, in_channels=1)
    synth.load_synth(net, 1337)
    net.to(dev).train()
    _eval_dropout_only(net)   # identical arithmetic in every run (Philox masks differ between draws)
    return net


def _loss(net, x, y, K, weights):
    from icl_b200.utils import losses as L
    o = net(x[:2], x[2:])
    ce, dice = L.seg_ce_dice(o[0], y[:2])
    return (weights[0] * dice + weights[1] * ce + weights[2] * L.AuxLoss3D(K)(o[2], y[:2]) + weights[3] * L.PseudoSoftLoss3D(K)(o[3], o[1])
            + weights[4] * L.softmax_mse_loss(o[3], o[4]))


def _worker(rank, world, port, K, use_graph, q):
    try:
        sys.path.insert(0, ROOT)
        import torch.distributed as dist
        from icl_b200 import parallel
        from icl_b200.optim import SGD
        from icl_b200.utils import synth
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        weights = (1.0, 1.0, 1.0, 0.1 if K == 16 else 1.0, 10.0)
        xs = [synth.synth_volume((4, 1, 96, 96, 96), 1338 + r).to(dev) for r in range(world)]
        ys = [synth.synth_labels((4, 96, 96, 96), K, 1339 + r).to(dev) for r in range(world)]
        steps = 2

        # ---- data-parallel run: fused-factored optimizer, overlapped exchange (optionally replayed as a CUDA graph)
        net = _build(K, dev)
        opt = SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4, fused_factored=True)
        dp = parallel.GradAverager(net, world, overlap=True)

        def dp_step(x, y):
            dp.begin_step()
            loss = _loss(net, x, y, K, weights)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            dp.average()
            opt.step()
            return loss

        if use_graph:
            from icl_b200.graph import GraphedStep
            snap = {k: v.detach().clone() for k, v in net.state_dict().items()}
            dp_step(xs[rank], ys[rank])                    # eager step: learns the per-bucket accumulation counts, creates momentum
            dp.check_consistent()
            g = GraphedStep(dp_step, (xs[rank], ys[rank]), opt, warmup=1)
            with torch.no_grad():                          # warm-up + capture advanced the state: rewind it
                for k, v in net.state_dict().items():
                    v.copy_(snap[k])
                for p in opt.state:
                    opt.state[p]["momentum_buffer"].zero_()
            for _ in range(steps):
                g(xs[rank], ys[rank])
            launched = None
            g.release()
        else:
            launched = []
            for i in range(steps):
                dp_step(xs[rank], ys[rank])
                if i == 0:
                    dp.check_consistent()
                launched.append(dp.launched_in_backward)
        torch.cuda.synchronize()
        dp_params = {k: p.detach() for k, p in net.named_parameters()}
        # both ranks must hold the same parameters
        for k in ("conv1.conv2.0.weight", "sspa.class_decoders.2.mlp2.fc1.weight", "uscl.class_decoders.2.mlp2.fc2.weight", "final.bias"):
            t = dp_params[k].double().sum().reshape(1)
            both = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(both, t)
            assert all(torch.equal(b, both[0]) for b in both), "ranks diverged on %s: %s" % (k, [float(b) for b in both])
        dp.close()

        result = {"launched": launched}
        if rank == 0:
            # ---- single-process comparator: mean of the rank-local gradients (materialised .grad), same SGD rule
            ref = _build(K, dev)
            init = {k: p.detach().clone() for k, p in ref.named_parameters()}
            opt_r = SGD(ref.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4, fused_factored=False)
            for _ in range(steps):
                acc = None
                for r in range(world):
                    opt_r.zero_grad(set_to_none=True)
                    _loss(ref, xs[r], ys[r], K, weights).backward()
                    gr = {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in ref.named_parameters()}
                    if acc is None:
                        acc = gr
                    else:
                        for k in acc:
                            if acc[k] is not None:
                                acc[k] += gr[k]
                for k, p in ref.named_parameters():
                    p.grad = None if acc[k] is None else acc[k] / world
                opt_r.step()
            torch.cuda.synchronize()
            worst = []
            for k, p in ref.named_parameters():
                d = (dp_params[k] - p.detach()).double().norm().item()
                upd = (p.detach() - init[k]).double().norm().item()
                if upd == 0.0:   # a parameter the reference never updates (grad is None) must not move under DP either
                    assert d == 0.0, k
                    continue
                if k.endswith(".0.bias") and not (k.startswith("sspa") or k.startswith("uscl")):
                    continue     # conv bias in front of InstanceNorm: mathematically-zero gradient, the "update" is rounding noise
                worst.append((d / upd, k))
            worst.sort(reverse=True)
            result["worst"] = worst[:5]
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok", result))
    except Exception:
        q.put((rank, "error", traceback.format_exc()))


@pytest.mark.parametrize("K,use_graph", [(2, False), (2, True), (16, False)])
def test_dp2_step_equals_mean_gradient_step(K, use_graph):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, K, use_graph, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = {}
    for _ in procs:
        rank, status, payload = q.get(timeout=900)
        out[rank] = (status, payload)
    for p in procs:
        p.join(timeout=120)
    for r, (status, payload) in out.items():
        assert status == "ok", "rank %d:\n%s" % (r, payload)
    worst = out[0][1]["worst"]
    # |p_dp - p_ref| relative to the size of the 2-step UPDATE |p_ref - p_init| of each tensor.  The bound is the RUN-TO-RUN noise of
    # this network, not the exchange: two single-process evaluations of the same batch already differ by up to ~1e-2 in the deep
    # layers' weight gradients (tools/dp_debug.py, profiles/r02g_dp_debug.log: "noise floor") because the InstanceNorm statistics
    # are accumulated with atomics, a handful of pre-activations within 1e-7 of zero flip their ReLU derivative, and at random
    # init a weight gradient is a heavily cancelling sum in which ONE voxel weighs ~1/sqrt(#voxels) (1.7 % at the 12^3 level).
    # An exchange bug (missing bucket, wrong 1/world, stale factors) shows up as an error of order 0.5 - 1.
    assert worst[0][0] <= 5e-2, worst
    if not use_graph:
        # the ICL-head bucket must have started from a grad-ready hook inside backward from the second step on
        assert out[0][1]["launched"][1] >= 1, out[0][1]["launched"]
