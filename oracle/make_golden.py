"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (build container only).

    python -m oracle.make_golden [--only NAME]

The reference has no golden vectors of its own, so these fixtures are the
pin: reference classes from /root/reference/code (oracle/ref_import.py), parameters and
inputs from oracle/synth.py (regenerable anywhere from shapes + seed), outputs stored here.
Large tensors are stored as checksums plus values at seeded sample positions
(``sample_idx``), small ones in full.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_import, synth  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def sample_idx(numel, n=64, seed=4242):
    rng = np.random.Generator(np.random.PCG64([seed, numel % (2 ** 31)]))
    return rng.integers(0, numel, size=min(n, numel), dtype=np.int64)


def summarize(t, n=64):
    """[sum, l2, sampled values...] in float64/float32 for a tensor (or zeros for None)."""
    t = t.detach().double().reshape(-1)
    idx = sample_idx(t.numel(), n)
    return np.array([t.sum().item(), t.norm().item()], dtype=np.float64), t[idx].float().numpy()


def pack_labels(labels, K):
    """Label map (ints < K) -> bit planes packed with np.packbits: uint8 [nbits, ceil(numel / 8)] (per-voxel argmax parity)."""
    a = np.asarray(labels).reshape(-1).astype(np.uint8)
    nbits = max(1, int(np.ceil(np.log2(K))))
    return np.stack([np.packbits((a >> b) & 1) for b in range(nbits)])


def unpack_labels(packed, numel):
    out = np.zeros(numel, dtype=np.int64)
    for b in range(packed.shape[0]):
        out |= np.unpackbits(packed[b])[:numel].astype(np.int64) << b
    return out


def eval_dropout_only(model):
    """Dropout / DropPath -> identity; BatchNorm stays in train mode."""
    for m in model.modules():
        if m.__class__.__name__ in ("Dropout", "DropPath"):
            m.eval()


def gen_unet3d(ns, name, fs, K, cin, size, seed, B=1):
    m = ns.unet_3D(feature_scale=fs, n_classes=K, in_channels=cin)
    synth.load_synth(m, seed)
    m.train()
    eval_dropout_only(m)
    x = synth.synth_volume((B, cin) + (size,) * 3, seed + 1)
    y = synth.synth_labels((B,) + (size,) * 3, K, seed + 2)
    logits = m(x)
    soft = torch.softmax(logits, 1)
    loss = torch.nn.CrossEntropyLoss()(logits, y) + ns.losses.DiceLoss(K)(soft, y.unsqueeze(1))
    loss.backward()
    out = dict(meta=np.array([fs, K, cin, size, seed, B]), logits=logits.detach().numpy(), loss=np.float64(loss.item()))
    for k, p in m.named_parameters():
        s, v = summarize(p.grad)
        out["gsum/" + k] = s
        out["gval/" + k] = v
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "loss", loss.item())


MINI = dict(in_chans=(64, 32, 16), res=[2, 4, 8], heads=(4, 2, 1), K=3, B=2)


def gen_icl_head(ns):
    c = MINI
    ic = ns.InherentConsistent(in_chans=c["in_chans"], depths=(2, 2, 2), patch_size=(2, 2, 2),
                               input_resolution=c["res"], num_classes=c["K"], num_heads=c["heads"])
    synth.load_synth(ic, 11)
    ic.train()
    eval_dropout_only(ic)
    feats = [synth.synth_volume((c["B"], ch) + (r,) * 3, 20 + i).requires_grad_(True)
             for i, (ch, r) in enumerate(zip(c["in_chans"], c["res"]))]
    fm_l, q_l = ic(feats, None, "labeled")
    fm_u, q_u = ic(feats, [q.detach() for q in q_l], "unlabeled")
    loss = sum((f ** 2).mean() for f in fm_l) + sum((f ** 2).mean() for f in fm_u) + sum((q ** 2).mean() for q in q_l)
    loss.backward()
    out = dict(loss=np.float64(loss.item()))
    for i in range(3):
        out["fm_l%d" % i] = fm_l[i].detach().numpy()
        out["fm_u%d" % i] = fm_u[i].detach().numpy()
        out["q_l%d" % i] = q_l[i].detach().numpy()
        out["dfeat%d" % i] = feats[i].grad.numpy()
    none = []
    for k, p in ic.named_parameters():
        if p.grad is None:
            none.append(k)
        elif p.numel() <= 70000:
            out["g/" + k] = p.grad.numpy()
        else:
            s, v = summarize(p.grad)
            out["gsum/" + k] = s
            out["gval/" + k] = v
    for k, v in ic.state_dict().items():
        if "running" in k:
            out["stat/" + k] = v.numpy()
    out["grad_none"] = np.array(none)
    np.savez_compressed(os.path.join(GOLDEN, "icl_head_mini.npz"), **out)
    print("icl_head_mini loss", loss.item(), "grad None:", none)


def gen_losses(ns, K, name):
    L = ns.losses
    B = 2
    labels = synth.synth_blobs((B, 96, 96, 96), K, 31)
    final_lab = synth.synth_volume((B, K, 96, 96, 96), 32).requires_grad_(True)
    final_unlab = synth.synth_volume((B, K, 96, 96, 96), 33)
    fms = [synth.synth_volume((B, K, r, r, r), 40 + i).mul_(2.0).requires_grad_(True) for i, r in enumerate((6, 12, 24))]
    fms2 = [synth.synth_volume((B, K, r, r, r), 50 + i).mul_(2.0).requires_grad_(True) for i, r in enumerate((6, 12, 24))]
    fms3 = [synth.synth_volume((B, K, r, r, r), 60 + i).mul_(2.0) for i, r in enumerate((6, 12, 24))]
    ce = torch.nn.CrossEntropyLoss()(final_lab, labels)
    dice = L.DiceLoss(K)(torch.softmax(final_lab, 1), labels.unsqueeze(1))
    aux = L.AuxLoss3D(K)(fms, labels)
    pse = L.PseudoSoftLoss3D(K)(fms2, final_unlab)
    cons = L.softmax_mse_loss(fms2, fms3)
    total = dice + ce + aux + pse + 10 * cons
    total.backward()
    out = dict(K=np.int64(K), ce=ce.item(), dice=dice.item(), aux=aux.item(), pse=pse.item(), cons=cons.item(),
               total=total.item())
    s, v = summarize(final_lab.grad, 256)
    out["dfinal_sum"], out["dfinal_val"] = s, v
    for i in range(3):
        out["daux%d" % i] = fms[i].grad.numpy()
        out["dpse%d" % i] = fms2[i].grad.numpy()
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, {k: out[k] for k in ("ce", "dice", "aux", "pse", "cons", "total")})


def gen_step(ns, K, name, weights):
    t0 = time.time()
    m = ns.unet_3D_icl(feature_scale=4, n_classes=K, in_channels=1)
    synth.load_synth(m, 1337)
    m.train()
    eval_dropout_only(m)
    x = synth.synth_volume((4, 1, 96, 96, 96), 1338)
    y = synth.synth_labels((4, 96, 96, 96), K, 1339)
    L = ns.losses
    o = m(x[:2], x[2:])
    soft = torch.softmax(o[0], 1)
    ce = torch.nn.CrossEntropyLoss()(o[0], y[:2])
    dice = L.DiceLoss(K)(soft, y[:2].unsqueeze(1))
    aux = L.AuxLoss3D(K)(o[2], y[:2])
    pse = L.PseudoSoftLoss3D(K)(o[3], o[1])
    cons = L.softmax_mse_loss(o[3], o[4])
    total = weights[0] * dice + weights[1] * ce + weights[2] * aux + weights[3] * pse + weights[4] * cons
    total.backward()
    out = dict(K=np.int64(K), weights=np.array(weights, dtype=np.float64), ce=ce.item(), dice=dice.item(),
               aux=aux.item(), pse=pse.item(), cons=cons.item(), total=total.item())
    for nm, t in (("final_lab", o[0]), ("final_unlab", o[1])):
        s, v = summarize(t, 4096)
        out[nm + "_sum"], out[nm + "_val"] = s, v
        out[nm + "_argmax_count"] = np.bincount(t.argmax(1).reshape(-1).numpy(), minlength=K)
        out[nm + "_argmax_bits"] = pack_labels(t.argmax(1).numpy(), K)
    for j, nm in ((2, "maps_lab"), (3, "maps_unlab"), (4, "maps_consis")):
        for i in range(3):
            t = o[j][i].detach()
            if t.numel() <= 70000:
                out["%s%d" % (nm, i)] = t.numpy()
            else:
                s, v = summarize(t, 4096)
                out["%s%d_sum" % (nm, i)], out["%s%d_val" % (nm, i)] = s, v
    none = []
    for k, p in m.named_parameters():
        if p.grad is None:
            none.append(k)
            continue
        s, v = summarize(p.grad)
        out["gsum/" + k] = s
        out["gval/" + k] = v
    out["grad_none"] = np.array(none)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "losses", {k: out[k] for k in ("ce", "dice", "aux", "pse", "cons", "total")},
          "none", len(none), "time %.1fs" % (time.time() - t0))


MINI2D = dict(in_chans=(64, 32, 16), res=[4, 8, 16], heads=(4, 2, 1), K=3, B=2)


def gen_unet2d(ns, name, K, size, seed, B):
    """2D UNet (config 1 backbone): logits, CE + Dice(softmax=True), parameter gradients, BatchNorm running statistics."""
    m = ns.UNet(in_chns=1, class_num=K)
    synth.load_synth(m, seed)
    m.train()
    eval_dropout_only(m)
    x = synth.synth_volume((B, 1, size, size), seed + 1)
    y = synth.synth_labels((B, size, size), K, seed + 2)
    logits = m(x)
    loss = torch.nn.CrossEntropyLoss()(logits, y) + ns.losses.DiceLoss(K)(logits, y.unsqueeze(1), softmax=True)
    loss.backward()
    out = dict(meta=np.array([K, size, seed, B]), logits=logits.detach().numpy(), loss=np.float64(loss.item()))
    for k, p in m.named_parameters():
        sm, v = summarize(p.grad)
        out["gsum/" + k] = sm
        out["gval/" + k] = v
    for k, v in m.state_dict().items():
        if "running" in k:
            out["stat/" + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "loss", loss.item())


def gen_icl_head2d(ns):
    c = MINI2D
    ic = ns.InherentConsistent2d(in_chans=c["in_chans"], depths=(2, 2, 2), patch_size=(2, 2), input_resolution=c["res"],
                                 num_classes=c["K"], num_heads=c["heads"])
    synth.load_synth(ic, 31)
    ic.train()
    eval_dropout_only(ic)
    feats = [synth.synth_volume((c["B"], ch, r, r), 40 + i).requires_grad_(True) for i, (ch, r) in enumerate(zip(c["in_chans"], c["res"]))]
    fm_l, q_l = ic(feats, None, "labeled")
    fm_u, q_u = ic(feats, [q.detach() for q in q_l], "unlabeled")
    loss = sum((f ** 2).mean() for f in fm_l) + sum((f ** 2).mean() for f in fm_u) + sum((q ** 2).mean() for q in q_l)
    loss.backward()
    out = dict(loss=np.float64(loss.item()))
    for i in range(3):
        out["fm_l%d" % i] = fm_l[i].detach().numpy()
        out["fm_u%d" % i] = fm_u[i].detach().numpy()
        out["q_l%d" % i] = q_l[i].detach().numpy()
        out["dfeat%d" % i] = feats[i].grad.numpy()
    none = []
    for k, p in ic.named_parameters():
        if p.grad is None:
            none.append(k)
        else:
            out["g/" + k] = p.grad.numpy()
    out["grad_none"] = np.array(none)
    np.savez_compressed(os.path.join(GOLDEN, "icl_head2d_mini.npz"), **out)
    print("icl_head2d_mini loss", loss.item(), "grad None:", len(none))


def gen_losses2d(ns, K=4, S=64, B=3):
    L = ns.losses
    labels = synth.synth_labels((B, S, S), K, 51)
    out_lab = synth.synth_volume((B, K, S, S), 52).requires_grad_(True)
    out_unlab = synth.synth_volume((B, K, S, S), 53)
    fms = [synth.synth_volume((B, K, r, r), 60 + i).mul_(2.0).requires_grad_(True) for i, r in enumerate((8, 16, 32))]
    fms2 = [synth.synth_volume((B, K, r, r), 70 + i).mul_(2.0).requires_grad_(True) for i, r in enumerate((8, 16, 32))]
    fms3 = [synth.synth_volume((B, K, r, r), 80 + i).mul_(2.0) for i, r in enumerate((8, 16, 32))]
    ce = torch.nn.CrossEntropyLoss()(out_lab, labels)
    dice = L.DiceLoss(K)(out_lab, labels.unsqueeze(1), softmax=True)
    aux = L.AuxLoss(K, resize=[S, S])(fms, labels)
    pse = L.PseudoSoftLoss(K, resize=[S, S])(fms2, out_unlab)
    cons = L.softmax_mse_loss(fms2, fms3)
    total = ce + dice + aux + pse + 50 * cons
    total.backward()
    out = dict(meta=np.array([K, S, B]), ce=ce.item(), dice=dice.item(), aux=aux.item(), pse=pse.item(), cons=cons.item(), total=total.item(),
               dout=out_lab.grad.numpy())
    for i in range(3):
        out["daux%d" % i] = fms[i].grad.numpy()
        out["dpse%d" % i] = fms2[i].grad.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "losses2d_k4.npz"), **out)
    print("losses2d_k4", {k: out[k] for k in ("ce", "dice", "aux", "pse", "cons", "total")})


def gen_step2d(ns):
    """Config 1: UNet_icl(1, 4), batch 24 (12 labeled + 12 unlabeled) x 1x256x256, one full forward + 5 losses + backward."""
    t0 = time.time()
    K = 4
    m = ns.UNet_icl(1, K)
    synth.load_synth(m, 1337)
    m.train()
    eval_dropout_only(m)
    x = synth.synth_volume((24, 1, 256, 256), 1338)
    y = synth.synth_labels((24, 256, 256), K, 1339)
    L = ns.losses
    o = m(x[:12], x[12:])
    ce = torch.nn.CrossEntropyLoss()(o[0], y[:12].long())
    dice = L.DiceLoss(K)(o[0], y[:12].unsqueeze(1), softmax=True)
    aux = L.AuxLoss(K, resize=[256, 256])(o[2], y[:12])
    pse = L.PseudoSoftLoss(K, resize=[256, 256])(o[3], o[1])
    cons = L.softmax_mse_loss(o[3], o[4])
    total = ce + dice + aux + pse + 50 * cons
    total.backward()
    out = dict(K=np.int64(K), ce=ce.item(), dice=dice.item(), aux=aux.item(), pse=pse.item(), cons=cons.item(), total=total.item())
    for nm, t in (("out_lab", o[0]), ("out_unlab", o[1])):
        sm, v = summarize(t, 4096)
        out[nm + "_sum"], out[nm + "_val"] = sm, v
        out[nm + "_argmax_count"] = np.bincount(t.argmax(1).reshape(-1).numpy(), minlength=K)
        out[nm + "_argmax_bits"] = pack_labels(t.argmax(1).numpy(), K)
    for j, nm in ((2, "maps_lab"), (3, "maps_unlab"), (4, "maps_consis")):
        for i in range(3):
            sm, v = summarize(o[j][i].detach(), 4096)
            out["%s%d_sum" % (nm, i)], out["%s%d_val" % (nm, i)] = sm, v
    none = []
    for k, p in m.named_parameters():
        if p.grad is None:
            none.append(k)
            continue
        sm, v = summarize(p.grad)
        out["gsum/" + k] = sm
        out["gval/" + k] = v
    out["grad_none"] = np.array(none)
    np.savez_compressed(os.path.join(GOLDEN, "step_cfg1.npz"), **out)
    print("step_cfg1 losses", {k: out[k] for k in ("ce", "dice", "aux", "pse", "cons", "total")}, "none", len(none), "time %.1fs" % (time.time() - t0))


def swin_config():
    """Effective values of configs/swin_tiny_patch4_window7_224_lite.yaml over networks/config.py defaults."""
    from types import SimpleNamespace as NS
    return NS(DATA=NS(IMG_SIZE=224), MODEL=NS(DROP_RATE=0.0, DROP_PATH_RATE=0.2, PRETRAIN_CKPT=None,
              SWIN=NS(PATCH_SIZE=4, IN_CHANS=3, EMBED_DIM=96, DEPTHS=[2, 2, 2, 2], NUM_HEADS=[3, 6, 12, 24], WINDOW_SIZE=7, MLP_RATIO=4.0,
                      QKV_BIAS=True, QK_SCALE=False, APE=False, PATCH_NORM=True)), TRAIN=NS(USE_CHECKPOINT=False))


def gen_step_swin(ns, name, n_lab, n_unlab, seed):
    """Config 4: SwinUnet(cfg, 224, 4) with the ICL heads, n_lab labeled + n_unlab unlabeled 1x224x224 slices, forward + the five
    losses of train_inherent_consistent_swinunet_2D.py:148-155 + backward."""
    t0 = time.time()
    K = 4
    m = ns.SwinUnet(swin_config(), img_size=224, num_classes=K)
    synth.load_synth(m, seed)
    m.train()
    eval_dropout_only(m)
    n = n_lab + n_unlab
    x = synth.synth_volume((n, 1, 224, 224), seed + 1)
    y = synth.synth_labels((n, 224, 224), K, seed + 2)
    L = ns.losses
    o = m(x[:n_lab], x[n_lab:])
    ce = torch.nn.CrossEntropyLoss()(o[0], y[:n_lab].long())
    dice = L.DiceLoss(K)(o[0], y[:n_lab].unsqueeze(1), softmax=True)
    aux = L.AuxLoss(K)(o[2], y[:n_lab])
    pse = L.PseudoSoftLoss(K)(o[3], o[1])
    cons = L.softmax_mse_loss(o[3], o[4])
    total = ce + dice + aux + pse + 50 * cons
    total.backward()
    out = dict(K=np.int64(K), n_lab=np.int64(n_lab), n_unlab=np.int64(n_unlab), seed=np.int64(seed), ce=ce.item(), dice=dice.item(),
               aux=aux.item(), pse=pse.item(), cons=cons.item(), total=total.item())
    for nm, t in (("out_lab", o[0]), ("out_unlab", o[1])):
        sm, v = summarize(t.detach(), 4096)
        out[nm + "_sum"], out[nm + "_val"] = sm, v
        out[nm + "_argmax_count"] = np.bincount(t.argmax(1).reshape(-1).numpy(), minlength=K)
        out[nm + "_argmax_bits"] = pack_labels(t.argmax(1).numpy(), K)
    for j, nm in ((2, "maps_lab"), (3, "maps_unlab"), (4, "maps_consis")):
        for i in range(3):
            sm, v = summarize(o[j][i].detach(), 4096)
            out["%s%d_sum" % (nm, i)], out["%s%d_val" % (nm, i)] = sm, v
    none = []
    for k, p in m.named_parameters():
        if p.grad is None:
            none.append(k)
            continue
        sm, v = summarize(p.grad)
        out["gsum/" + k] = sm
        out["gval/" + k] = v
    out["grad_none"] = np.array(none)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "losses", {k: out[k] for k in ("ce", "dice", "aux", "pse", "cons", "total")}, "none", len(none), "time %.1fs" % (time.time() - t0))


def gen_state_keys_2d_swin(ns):
    """state_dict / parameter name+shape lists of the 2D and Swin models, merged into tests/golden/state_keys.json."""
    import json
    path = os.path.join(GOLDEN, "state_keys.json")
    keys = json.load(open(path)) if os.path.exists(path) else {}
    for name, m in (("unet_icl_k4", ns.UNet_icl(1, 4)), ("unet2d_k4", ns.UNet(1, 4)),
                    ("swin_unet_k4", ns.SwinUnet(swin_config(), img_size=224, num_classes=4))):
        keys[name] = [[k, list(v.shape)] for k, v in m.state_dict().items()]
        keys[name + "_params"] = [k for k, _ in m.named_parameters()]
    json.dump(keys, open(path, "w"))
    print("state_keys:", sorted(keys.keys()))


def gen_sliding(ns):
    tsc = ref_import.load_test_single_case()
    m = ns.unet_3D(feature_scale=4, n_classes=2, in_channels=1)
    synth.load_synth(m, 77)
    m.eval()
    image = synth.synth_volume((120, 104, 90), 78).numpy()
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self  # the reference calls .cuda() on each patch (:123)
    try:
        label = tsc(m, image, 64, 64, (96, 96, 96), num_classes=2)
    finally:
        torch.Tensor.cuda = orig_cuda
    gt = synth.synth_blobs((120, 104, 90), 2, 79).numpy()
    pred = label.astype(np.int64)
    inter = int(np.count_nonzero((pred > 0) & (gt > 0)))
    out = dict(shape=np.array(image.shape), label_bits=np.packbits(label.astype(np.uint8).reshape(-1)),
               counts=np.array([inter, int((pred > 0).sum()), int((gt > 0).sum())], dtype=np.int64),
               dice=np.float64(2.0 * inter / float((pred > 0).sum() + (gt > 0).sum())))
    np.savez_compressed(os.path.join(GOLDEN, "sliding_window.npz"), **out)
    print("sliding_window label fraction", label.mean(), "dice", out["dice"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ns = ref_import.load()
    jobs = {
        "unet3d_fs4": lambda: gen_unet3d(ns, "unet3d_fs4_k2_c1_32", 4, 2, 1, 32, 101),
        "unet3d_fs16": lambda: gen_unet3d(ns, "unet3d_fs16_k3_c2_32", 16, 3, 2, 32, 202, B=2),
        "icl_head": lambda: gen_icl_head(ns),
        "losses_k2": lambda: gen_losses(ns, 2, "losses_k2"),
        "losses_k5": lambda: gen_losses(ns, 5, "losses_k5"),
        "step_cfg2": lambda: gen_step(ns, 2, "step_cfg2", (1, 1, 1, 1, 10)),
        "step_cfg3": lambda: gen_step(ns, 16, "step_cfg3", (1, 1, 1, 0.1, 10)),
        "sliding": lambda: gen_sliding(ns),
        "unet2d": lambda: gen_unet2d(ns, "unet2d_k4_64", 4, 64, 21, 4),
        "icl_head2d": lambda: gen_icl_head2d(ns),
        "losses2d": lambda: gen_losses2d(ns),
        "step_cfg1": lambda: gen_step2d(ns),
        "state_keys_2d_swin": lambda: gen_state_keys_2d_swin(ns),
        "step_swin_b2": lambda: gen_step_swin(ns, "step_swin_b2", 1, 1, 4401),
        "step_cfg4": lambda: gen_step_swin(ns, "step_cfg4", 8, 8, 4404),
    }
    for k, fn in jobs.items():
        if a.only is None or a.only == k:
            fn()


if __name__ == "__main__":
    main()
