"""Stand-in for monai.networks.layers (MONAI 1.0.1): Conv factory, DropPath, trunc_normal_.

Semantics restated from MONAI/timm: ``Conv[Conv.CONV, d]`` is ``torch.nn.Conv{d}d``;
``DropPath`` draws one Bernoulli(keep) per sample (dim 0) in training and scales by
1/keep, and is the identity in eval mode or when drop_prob == 0.
"""
import torch.nn as nn
from torch.nn.init import trunc_normal_  # noqa: F401


class _ConvFactory:
    CONV = "conv"
    CONVTRANS = "convtrans"

    def __getitem__(self, key):
        name, dim = key
        table = {("conv", 1): nn.Conv1d, ("conv", 2): nn.Conv2d, ("conv", 3): nn.Conv3d}
        return table[(name, dim)]


Conv = _ConvFactory()


class DropPath(nn.Module):
    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        r = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            r.div_(keep)
        return x * r
