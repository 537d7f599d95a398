"""Localise 2D UNet gradient differences: icl_b200 UNet vs a plain-torch (fp32, TF32 off) replica using the same parameters,
listed per parameter from the output backwards, run twice to expose run-to-run variation."""
import copy
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from icl_b200.networks.unet import UNet  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def block(cb, x):
    s = cb.conv_conv
    y = F.leaky_relu(F.batch_norm(F.conv2d(x, s[0].weight, s[0].bias, padding=1), None, None, s[1].weight, s[1].bias, True), 0.01)
    return F.leaky_relu(F.batch_norm(F.conv2d(y, s[4].weight, s[4].bias, padding=1), None, None, s[5].weight, s[5].bias, True), 0.01)


def ref_forward(net, x):
    e, d = net.encoder, net.decoder
    x0 = block(e.in_conv, x)
    feats = [x0]
    for dn in (e.down1, e.down2, e.down3, e.down4):
        feats.append(block(dn.maxpool_conv[1], F.max_pool2d(feats[-1], 2)))
    y = feats[4]
    for up, skip in ((d.up1, feats[3]), (d.up2, feats[2]), (d.up3, feats[1]), (d.up4, feats[0])):
        y1 = F.interpolate(F.conv2d(y, up.conv1x1.weight, up.conv1x1.bias), scale_factor=2, mode="bilinear", align_corners=True)
        y = block(up.conv, torch.cat([skip, y1], 1))
    return F.conv2d(y, d.out_conv.weight, d.out_conv.bias, padding=1)


def run(size, B, K=4, dbl=False):
    torch.manual_seed(7)
    net = UNet(1, K)  # constructor init; BatchNorm affine parameters perturbed so that gamma / beta gradients are exercised
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.3, 0.3)
    net.cuda().train()
    for m in net.modules():
        if m.__class__.__name__ == "Dropout":
            m.eval()
    ref = copy.deepcopy(net)
    if dbl:
        ref = ref.double()
    g = torch.Generator().manual_seed(8)
    x = torch.randn(B, 1, size, size, generator=g).cuda()
    y = torch.randint(0, K, (B, size, size), generator=g).cuda()
    lr = ref_forward(ref, x.double() if dbl else x)
    (F.cross_entropy(lr, y.long())).backward()
    outs = []
    for rep in range(2):
        for p in net.parameters():
            p.grad = None
        lo = net(x)
        F.cross_entropy(lo, y.long()).backward()
        outs.append({k: p.grad.clone() for k, p in net.named_parameters()})
    print("size %d B %d  logits rel %.2e" % (size, B, ((lo - lr).norm() / lr.norm()).item()))
    names = [k for k, _ in net.named_parameters()][::-1]
    rg = dict(ref.named_parameters())
    for k in names:
        g = rg[k].grad.float()
        n = g.norm().item()
        e0 = (outs[0][k] - g).norm().item()
        e1 = (outs[1][k] - g).norm().item()
        rr = (outs[0][k] - outs[1][k]).norm().item()
        flag = " <<<" if e0 > 1e-3 * n + 2e-6 else ""
        print("  %-48s |g| %.3e  err %.2e %.2e  run2run %.2e%s" % (k, n, e0 / (n + 1e-30), e1 / (n + 1e-30), rr / (n + 1e-30), flag))


if __name__ == "__main__":
    import icl_b200
    for mode in sys.argv[1:] or ["parity"]:
        icl_b200.set_precision(mode)
        print("=== precision", mode)
        run(64, 4, dbl=True)
        run(128, 2, dbl=True)
