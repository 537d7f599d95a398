"""Parameter-holding blocks with the reference's names (networks/utils.py:99-123, 260-276).

These modules only OWN parameters (so state_dict()/load_state_dict() round-trip with the reference in both
directions, SURVEY.md §5 checkpoint row); the arithmetic runs in backbone3d.Backbone3DFn.  The nn.Conv3d
children are never called.
"""
import torch.nn as nn


def _kaiming(conv):
    # weights_init_kaiming (networks_other.py:40-49): kaiming-normal fan_in on conv weights, bias untouched
    nn.init.kaiming_normal_(conv.weight.data, a=0, mode="fan_in")


class UnetConv3(nn.Module):
    """(Conv3d 3^3 p1 + bias -> InstanceNorm3d -> ReLU) x 2; is_batchnorm=True means InstanceNorm (utils.py:103-109)."""

    def __init__(self, in_size, out_size, is_batchnorm=True, kernel_size=(3, 3, 3), padding_size=(1, 1, 1), init_stride=(1, 1, 1)):
        super().__init__()
        if tuple(kernel_size) != (3, 3, 3) or tuple(padding_size) != (1, 1, 1) or tuple(init_stride) != (1, 1, 1) or not is_batchnorm:
            raise NotImplementedError("icl_b200 UnetConv3 implements the 3x3x3 / pad 1 / stride 1 / InstanceNorm configuration "
                                      "used by unet_3D and unet_3D_icl")
        self.conv1 = nn.Sequential(nn.Conv3d(in_size, out_size, 3, 1, 1))
        self.conv2 = nn.Sequential(nn.Conv3d(out_size, out_size, 3, 1, 1))
        _kaiming(self.conv1[0])
        _kaiming(self.conv2[0])

    def params(self):
        return [self.conv1[0].weight, self.conv1[0].bias, self.conv2[0].weight, self.conv2[0].bias]


class UnetUp3_CT(nn.Module):
    """Upsample(trilinear x2) + cat([skip, up]) + UnetConv3(in+out -> out) (utils.py:260-276)."""

    def __init__(self, in_size, out_size, is_batchnorm=True):
        super().__init__()
        self.conv = UnetConv3(in_size + out_size, out_size, is_batchnorm)

    def params(self):
        return self.conv.params()
