"""GPU parity tests of whole backbones against live CPU / float64 evaluations of the same arithmetic:
  * 3D backbone backward with the discrete choices (ReLU signs, MaxPool winners) held fixed: every gradient within 1e-3 of float64;
  * input sizes whose `center` level has an odd depth (48^3);
  * the inference=True early return of unet_3D_icl and the eval-mode (running statistics) forward of the 2D UNet / UNet_icl;
  * sliding-window inference on the BASELINE config-5 grid (240 x 240 x 155, 96^3 windows) at strides 64 (reference default,
    test_3D_BraTS.py:155) and 48 (50 % overlap), window-batched, against the oracle's restatement of test_single_case."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import assert_close
from oracle import restate as R
from oracle import restate2d as R2
from oracle import synth

pytestmark = pytest.mark.gpu


def eval_dropout_only(model):
    for m in model.modules():
        if m.__class__.__name__ in ("Dropout", "DropPath"):
            m.eval()


def _replica_backbone3d(P, x, acts):
    """unet_3D forward in plain torch on float64 parameters `P`; ReLU sign patterns come from `acts` (the post-ReLU activations of
    the icl_b200 forward, NDHWC, in call order) and MaxPool3d winners from those same activations.
    Returns (logits, discrete disagreements, elements)."""
    it = iter(acts)
    cnt = [0, 0]
    last = {}

    def half(prefix, t):
        pre = F.instance_norm(F.conv3d(t, P[prefix + ".0.weight"], P[prefix + ".0.bias"], padding=1), eps=1e-5)
        ours = next(it).permute(0, 4, 1, 2, 3)
        m = ours > 0
        cnt[0] += int(((pre.detach() > 0) != m).sum())
        cnt[1] += m.numel()
        last["a"] = ours
        return pre * m

    def block(prefix, t):
        return half(prefix + ".conv2", half(prefix + ".conv1", t))

    def pool(t, ours):
        idx = F.max_pool3d(ours, 2, return_indices=True)[1]
        cnt[0] += int((F.max_pool3d(t.detach(), 2, return_indices=True)[1] != idx).sum())
        cnt[1] += idx.numel()
        n, c, d, h, w = idx.shape
        return t.flatten(2).gather(2, idx.flatten(2)).view(n, c, d, h, w)

    feats, ours_feats = [], []
    t = x
    for name in ("conv1", "conv2", "conv3", "conv4"):
        t = block(name, t)
        feats.append(t)
        ours_feats.append(last["a"])
        t = pool(t, last["a"].double())
    t = block("center", t)
    for name, skip in (("up_concat4", 3), ("up_concat3", 2), ("up_concat2", 1), ("up_concat1", 0)):
        up = F.interpolate(t, scale_factor=(2, 2, 2), mode="trilinear", align_corners=False)
        t = block(name + ".conv", torch.cat([feats[skip], up], 1))
    return F.conv3d(t, P["final.weight"], P["final.bias"]), cnt[0], cnt[1]


@pytest.mark.parametrize("mode,size,B,K", [("parity", 32, 2, 2), ("parity", 64, 1, 16), ("fp32", 32, 2, 2)])
def test_backbone3d_grads_mask_matched(monkeypatch, mode, size, B, K):
    """Whole 3D backbone forward + backward (tcgen05 convolutions, InstanceNorm / pool / upsample kernels, fused head) against a
    float64 torch replica with the same parameters and the same ReLU signs / MaxPool winners: logits 5e-4, every parameter
    gradient 1e-3; the two forwards may disagree on at most 2e-5 of the discrete choices."""
    import icl_b200
    from icl_b200.networks import backbone3d
    from icl_b200.networks.unet_3D import unet_3D
    net = unet_3D(feature_scale=4, n_classes=K, in_channels=1)
    synth.load_synth(net, 501)
    net.cuda().train()
    eval_dropout_only(net)
    x = synth.synth_volume((B, 1, size, size, size), 502).cuda()
    y = synth.synth_labels((B, size, size, size), K, 503).cuda()
    acts = []
    orig = backbone3d.ops.instnorm_relu_fwd

    def recording(*a, **kw):
        out = orig(*a, **kw)
        if out[0] is not None:
            acts.append(out[0].detach())
        else:   # between the two convolutions of a block only the split-bf16 operand exists: hi + lo carries ~16 mantissa bits
            pk = out[1]
            B_, C8, D_, H_, W_ = pk.shape[1:6]
            acts.append(pk.float().sum(0).permute(0, 2, 3, 4, 1, 5).reshape(B_, D_, H_, W_, C8 * 8))
        return out

    monkeypatch.setattr(backbone3d.ops, "instnorm_relu_fwd", recording)
    icl_b200.set_precision(mode)
    try:
        logits = net(x)
        F.cross_entropy(logits, y).backward()
    finally:
        icl_b200.set_precision("parity")
    assert len(acts) == 18
    P = {k: v.detach().double().requires_grad_(True) for k, v in net.state_dict().items()}
    lr, nflip, nel = _replica_backbone3d(P, x.double(), acts)
    F.cross_entropy(lr, y).backward()
    assert nflip <= max(3, 2e-5 * nel), "%d of %d discrete choices differ" % (nflip, nel)
    assert_close(logits.detach().cpu(), lr.detach().cpu(), 5e-4, "logits")
    for k, p in net.named_parameters():
        if k.endswith(".0.bias"):   # conv bias in front of InstanceNorm: the exact gradient is 0
            assert p.grad.abs().max().item() < 1e-5, k
            continue
        assert_close(p.grad.cpu(), P[k].grad.cpu(), 1e-3, k, abs_floor=1e-8)


def test_backbone3d_odd_center_depth_48():
    """48^3 input: the `center` level is 3^3 (odd depth), where the tensor-core weight gradient does not apply — forward, backward
    and eval must all work (CUDA-core weight gradient for that level) and match the CPU oracle."""
    from icl_b200.networks.unet_3D import unet_3D
    net = unet_3D(feature_scale=4, n_classes=2, in_channels=1)
    synth.load_synth(net, 511)
    net.cuda().train()
    eval_dropout_only(net)
    x = synth.synth_volume((1, 1, 48, 48, 48), 512)
    y = synth.synth_labels((1, 48, 48, 48), 2, 513)
    logits = net(x.cuda())
    F.cross_entropy(logits, y.cuda()).backward()
    P = R.make_params({k: v.detach().cpu() for k, v in net.state_dict().items()})
    ref = R.unet_3d_forward(P, x)
    R.ce_loss(ref, y).backward()
    assert_close(logits.detach().cpu(), ref.detach(), 5e-4, "logits 48^3")
    for k in ("center.conv1.0.weight", "center.conv2.0.weight", "up_concat4.conv.conv1.0.weight", "conv1.conv2.0.weight", "final.weight"):
        # flip-sized sanity bound (the arithmetic is pinned by the mask-matched test): at the 3^3 `center` level one ReLU flip is 1/27
        # of a channel's statistics, so the bound is loose
        assert_close(dict(net.named_parameters())[k].grad.cpu(), P[k].grad, 6e-2, k)
    net.eval()
    with torch.no_grad():
        assert_close(net(x.cuda()).cpu(), ref.detach(), 5e-4, "eval logits 48^3")


@pytest.mark.parametrize("K", [2, 16])
def test_unet3d_icl_inference_flag(K):
    """unet_3D_icl(x, inference=True) returns the labeled-branch logits before any ICL head runs (unet_3D_icl.py:119-120), in train
    and eval mode; equals the CPU oracle and unet_3D with the same backbone weights."""
    from icl_b200.networks.unet_3D import unet_3D
    from icl_b200.networks.unet_3D_icl import unet_3D_icl
    net = unet_3D_icl(feature_scale=4, n_classes=K, in_channels=1)
    synth.load_synth(net, 521)
    net.cuda().eval()
    x = synth.synth_volume((1, 1, 96, 96, 96), 522)
    with torch.no_grad():
        out = net(x.cuda(), inference=True)
    assert tuple(out.shape) == (1, K, 96, 96, 96)
    P = R.make_params({k: v.detach().cpu() for k, v in net.state_dict().items()}, requires_grad=False)
    with torch.no_grad():
        ref = R.unet_3d_icl_forward(P, x, inference=True, training=False)
    assert_close(out.cpu(), ref, 5e-4, "inference=True logits")
    agree = float((out.argmax(1).cpu() == ref.argmax(1)).float().mean())
    assert agree >= 0.999, agree
    plain = unet_3D(feature_scale=4, n_classes=K, in_channels=1)
    plain.load_state_dict({k: v for k, v in net.state_dict().items() if not (k.startswith("sspa") or k.startswith("uscl"))})
    plain.cuda().eval()
    with torch.no_grad():
        assert_close(plain(x.cuda()).cpu(), out.cpu(), 1e-4, "unet_3D with the ICL checkpoint's backbone keys")   # statistics use atomics


def test_unet2d_eval_mode_running_stats():
    """2D UNet / UNet_icl in eval mode: BatchNorm uses the running statistics (functional2d.Conv2dBnActFn, training=False branch),
    Dropout is the identity; inference=True of UNet_icl returns the labeled logits only (unet_icl.py:237-246)."""
    from icl_b200.networks.unet import UNet
    from icl_b200.networks.unet_icl import UNet_icl
    K = 4
    x = synth.synth_volume((3, 1, 64, 64), 532)
    net = UNet(1, K)
    synth.load_synth(net, 531)    # running_mean / running_var are perturbed away from 0 / 1 by synth
    net.cuda().eval()
    with torch.no_grad():
        out = net(x.cuda())
    P = R.make_params({k: v.detach().cpu() for k, v in net.state_dict().items()}, requires_grad=False)
    with torch.no_grad():
        ref = R2.unet2d_forward(P, x, training=False)
    assert_close(out.cpu(), ref, 5e-4, "UNet eval logits")
    assert float((out.argmax(1).cpu() == ref.argmax(1)).float().mean()) >= 0.999
    before = {k: v.clone() for k, v in net.state_dict().items() if "running" in k or "num_batches" in k}
    with torch.no_grad():
        net(x.cuda())
    for k, v in net.state_dict().items():
        if k in before:
            assert torch.equal(v, before[k]), "eval forward must not touch " + k
    icl = UNet_icl(1, K)
    synth.load_synth(icl, 533)
    icl.cuda().eval()
    x2 = synth.synth_volume((2, 1, 256, 256), 534)
    with torch.no_grad():
        o2 = icl(x2.cuda(), inference=True)
    P2 = R.make_params({k: v.detach().cpu() for k, v in icl.state_dict().items()}, requires_grad=False)
    with torch.no_grad():
        r2 = R2.unet_icl_forward(P2, x2, inference=True, training=False)
    assert_close(o2.cpu(), r2, 5e-4, "UNet_icl eval inference=True logits")


@pytest.mark.parametrize("stride,window_batch", [(64, 4), (48, 4), (48, 1)])
def test_sliding_window_config5_grid(stride, window_batch):
    """BASELINE config 5 grid: one 240 x 240 x 155 volume, 96^3 windows.  Stride 64 is the reference's default
    (test_3D_BraTS.py:155: 4*4*2 = 32 windows), stride 48 the 50 % overlap BASELINE names (4*4*3 = 48 windows).  The device path
    (window-batched network calls, on-device accumulate / argmax) must agree with the oracle's restatement of test_single_case on
    >= 99.9 % of the voxels; the window grid and visit counts must be identical."""
    from icl_b200 import inference
    from icl_b200.networks.unet_3D import unet_3D
    shape, patch, K = (240, 240, 155), (96, 96, 96), 2
    net = unet_3D(feature_scale=4, n_classes=K, in_channels=1)
    synth.load_synth(net, 77)
    net.cuda().eval()
    image = synth.synth_volume(shape, 78).numpy()
    label = inference.test_single_case(net, image, stride, stride, patch, num_classes=K, window_batch=window_batch)
    assert label.shape == shape and label.dtype == np.int64
    score, cnt, _ = inference.sliding_window_scores(net, image, stride, stride, patch, K, window_batch=window_batch)
    # reference visit counts from the window grid of test_3D_BraTS.py:106-118
    want_cnt = np.zeros(shape, dtype=np.float32)
    starts = [inference.window_starts(s, p, stride) for s, p in zip(shape, patch)]
    assert [len(s) for s in starts] == ([4, 4, 2] if stride == 64 else [4, 4, 3])
    for xs in starts[0]:
        for ys in starts[1]:
            for zs in starts[2]:
                want_cnt[xs:xs + 96, ys:ys + 96, zs:zs + 96] += 1
    assert np.array_equal(cnt.cpu().numpy(), want_cnt)
    if window_batch == 4:
        # the CPU oracle takes ~0.4 s per window: check a sub-volume that covers window overlaps in all three axes
        sub = image[:144, :96, :155] if stride == 48 else image[:160, :96, :155]
        got = inference.test_single_case(net, sub, stride, stride, patch, num_classes=K, window_batch=window_batch)
        P = R.make_params({k: v.detach().cpu() for k, v in net.state_dict().items()}, requires_grad=False)
        with torch.no_grad():
            want = R.test_single_case(lambda p: R.unet_3d_forward(P, p), sub, stride, stride, patch, K)
        agree = float((got == want).mean())
        assert agree >= 0.999, agree
    else:
        # window batching must not change the result beyond fp32 re-association of nothing: the accumulation order is the same
        again = inference.test_single_case(net, image, stride, stride, patch, num_classes=K, window_batch=4)
        assert float((again == label).mean()) >= 0.99999


def test_validation_driver_matches_oracle():
    """inference.test_all_case (in-training validation, val_3D.py:100-118): per-case per-class Dice equals the oracle's
    restatement (CPU sliding window + MedPy-dc restatement) on the same label maps; counts bit-exact on identical maps."""
    from icl_b200 import inference
    from icl_b200.networks.unet_3D import unet_3D
    K, patch = 2, (96, 96, 96)
    net = unet_3D(feature_scale=4, n_classes=K, in_channels=1)
    synth.load_synth(net, 77)
    net.cuda().eval()
    cases = [(synth.synth_volume((112, 96, 100), 90 + i).numpy(), synth.synth_blobs((112, 96, 100), K, 95 + i).numpy()) for i in range(2)]
    got = inference.test_all_case(net, cases, num_classes=K, patch_size=patch, stride_xy=64, stride_z=64, window_batch=2)
    assert len(got) == K - 1 and len(got[0]) == len(cases)
    P = R.make_params({k: v.detach().cpu() for k, v in net.state_dict().items()}, requires_grad=False)
    for ci, (image, label) in enumerate(cases):
        with torch.no_grad():
            want = R.test_single_case(lambda p: R.unet_3d_forward(P, p), image, 64, 64, patch, K)
        pred = inference.test_single_case(net, image, 64, 64, patch, num_classes=K, window_batch=2)
        assert float((pred == want).mean()) >= 0.999
        d_ref, counts_ref = R.dice_metric((pred == 1), (label == 1))           # on OUR label map: counts must be bit-exact
        d_mine, counts = inference.dice_metric(torch.from_numpy((pred == 1).astype(np.int64)).cuda(), torch.from_numpy((label == 1).astype(np.int64)).cuda())
        assert tuple(counts) == tuple(counts_ref) and d_mine == d_ref
        assert abs(got[0][ci][0] - d_ref) < 1e-3   # two separate inference runs: the label maps may differ in a few voxels (statistics use atomics)
