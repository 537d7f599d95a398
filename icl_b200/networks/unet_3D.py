"""unet_3D — drop-in for the reference's networks/unet_3D.py (constructor :22, forward :71-94)."""
import torch
import torch.nn as nn

from .backbone3d import N_LOW, Backbone3DFn, BackbonePairFn, BackbonePairLowFn, BackbonePairTopFn, _PairState
from .utils import UnetConv3, UnetUp3_CT, _kaiming


class _Backbone3DModule(nn.Module):
    """Shared constructor/body of unet_3D and unet_3D_icl (identical backbone state_dict keys)."""

    def __init__(self, feature_scale=4, n_classes=21, is_deconv=True, in_channels=3, is_batchnorm=True):
        super().__init__()
        self.is_deconv, self.in_channels, self.is_batchnorm, self.feature_scale = is_deconv, in_channels, is_batchnorm, feature_scale
        filters = [int(x / feature_scale) for x in (64, 128, 256, 512, 1024)]
        self.filters = filters
        self.conv1 = UnetConv3(in_channels, filters[0], is_batchnorm)
        self.maxpool1 = nn.MaxPool3d(kernel_size=(2, 2, 2))
        self.conv2 = UnetConv3(filters[0], filters[1], is_batchnorm)
        self.maxpool2 = nn.MaxPool3d(kernel_size=(2, 2, 2))
        self.conv3 = UnetConv3(filters[1], filters[2], is_batchnorm)
        self.maxpool3 = nn.MaxPool3d(kernel_size=(2, 2, 2))
        self.conv4 = UnetConv3(filters[2], filters[3], is_batchnorm)
        self.maxpool4 = nn.MaxPool3d(kernel_size=(2, 2, 2))
        self.center = UnetConv3(filters[3], filters[4], is_batchnorm)
        self.up_concat4 = UnetUp3_CT(filters[4], filters[3], is_batchnorm)
        self.up_concat3 = UnetUp3_CT(filters[3], filters[2], is_batchnorm)
        self.up_concat2 = UnetUp3_CT(filters[2], filters[1], is_batchnorm)
        self.up_concat1 = UnetUp3_CT(filters[1], filters[0], is_batchnorm)
        self.final = nn.Conv3d(filters[0], n_classes, 1)
        _kaiming(self.final)
        self.dropout1 = nn.Dropout(p=0.3)
        self.dropout2 = nn.Dropout(p=0.3)
        # test hook: explicit keep-masks (uint8, NDHWC order) consumed two per backbone pass instead of Philox
        self._mask_queue = []

    def _backbone_params(self):
        ps = []
        for blk in (self.conv1, self.conv2, self.conv3, self.conv4, self.center, self.up_concat4, self.up_concat3, self.up_concat2,
                    self.up_concat1):
            ps += blk.params()
        return ps + [self.final.weight, self.final.bias]

    def _drop_cfg(self):
        """Dropout(0.3) after `center` and after `up1` (unet_3D_icl.py:110,116): identity unless the Dropout
        modules are in training mode.  Seeds come from torch's CUDA generator (deterministic under manual_seed)."""
        if not (self.dropout1.training or self.dropout2.training):
            return None
        if self._mask_queue:
            m1, m2 = self._mask_queue.pop(0), self._mask_queue.pop(0)
            return (self.dropout1.p, m1, m2, 0, 0)
        # seeds are drawn on the device (torch's CUDA generator: deterministic under manual_seed, no host sync, and the
        # draw itself is CUDA-graph safe, so every replay of a captured step gets fresh masks)
        s = torch.randint(0, 2 ** 62, (2,), dtype=torch.int64, device=self.final.weight.device)
        return (self.dropout1.p, None, None, s[0:1], s[1:2])

    def _run(self, x):
        if not x.is_cuda:
            raise RuntimeError("icl_b200 networks run on CUDA tensors only (no CPU fallback)")
        return Backbone3DFn.apply(x, self._drop_cfg(), *self._backbone_params())

    def _pair_cfg(self, x_lab, x_unlab):
        if not (x_lab.is_cuda and x_unlab.is_cuda):
            raise RuntimeError("icl_b200 networks run on CUDA tensors only (no CPU fallback)")
        cfg_l = self._drop_cfg()
        cfg = cfg_l
        if cfg_l is not None and cfg_l[1] is not None:
            # explicit keep-masks (test hook): two per pass in the reference's order (labeled pass first), joined along the batch
            cfg_u = self._drop_cfg()
            cfg = (cfg_l[0], torch.cat([cfg_l[1], cfg_u[1]]), torch.cat([cfg_l[2], cfg_u[2]]), 0, 0)
        return cfg

    def _run_pair(self, x_lab, x_unlab):
        """Labeled and unlabeled pass as one batched pass (backbone3d.BackbonePairFn).  Returns
        ((final, center, up4, up3) of the labeled samples, the same for the unlabeled samples)."""
        cfg = self._pair_cfg(x_lab, x_unlab)
        o = BackbonePairFn.apply(torch.cat([x_lab, x_unlab]), x_lab.shape[0], cfg, *self._backbone_params())
        return (o[0], o[2], o[4], o[6]), (o[1], o[3], o[5], o[7])

    def _run_pair_low(self, x_lab, x_unlab):
        """The batched pass up to up3 as an autograd node of its own (backbone3d.BackbonePairLowFn).  Returns
        ((center, up4, up3) of the labeled samples, the same for the unlabeled samples, handle for _run_pair_top)."""
        cfg = self._pair_cfg(x_lab, x_unlab)
        ps = self._backbone_params()
        st = _PairState()
        o = BackbonePairLowFn.apply(torch.cat([x_lab, x_unlab]), x_lab.shape[0], cfg, st, list(ps[N_LOW:N_LOW + 8:2]),
                                    *ps[:N_LOW])
        return (o[0], o[2], o[4]), (o[1], o[3], o[5]), (o[6], st, ps[N_LOW:])

    @staticmethod
    def _run_pair_top(handle):
        """up_concat2, up_concat1, final on top of _run_pair_low: (final_lab, final_unlab)."""
        link, st, ps = handle
        return BackbonePairTopFn.apply(link, st, *ps)


class unet_3D(_Backbone3DModule):
    def forward(self, inputs):
        return self._run(inputs)[0]

    @staticmethod
    def apply_argmax_softmax(pred):
        return torch.softmax(pred, dim=1)
