"""timm.models.layers stand-ins: DropPath (one Bernoulli(keep) per sample, scaled by 1/keep, identity in eval — timm's
drop_path semantics), to_2tuple, trunc_normal_ (torch.nn.init's, same truncated-normal definition)."""
import torch.nn as nn
from torch.nn.init import trunc_normal_  # noqa: F401


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


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
