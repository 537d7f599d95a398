"""2D U-Net — drop-in for the reference's networks/unet.py `UNet` (constructor :306, forward :318-321) and the blocks it
shares with networks/unet_icl.py (ConvBlock :41-57, DownBlock :60-72, UpBlock :75-96, Encoder :128-155, Decoder :157-194).

The module tree, parameter / buffer names and their order are the reference's (checkpoints interchange); the arithmetic
runs in icl_b200.functional2d.  The nn.Conv2d / nn.BatchNorm2d / nn.LeakyReLU children only own parameters."""
import torch
import torch.nn as nn

from .. import functional2d as F2


class ConvBlock(nn.Module):
    """two convolution layers with batch norm and leaky relu (Dropout(p) between them)."""

    def __init__(self, in_channels, out_channels, dropout_p):
        super().__init__()
        self.conv_conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1),
            nn.BatchNorm2d(out_channels),
            nn.LeakyReLU(),
            nn.Dropout(dropout_p),
            nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1),
            nn.BatchNorm2d(out_channels),
            nn.LeakyReLU())
        self._mask_queue = []  # test hook: explicit keep-masks (uint8, NHWC order) instead of Philox

    def forward(self, x, x_up=None):
        """x_up: optional second source, concatenated after x along channels (UpBlock's cat([x2, x1], 1))."""
        if not x.is_cuda:
            raise RuntimeError("icl_b200 networks run on CUDA tensors only (no CPU fallback)")
        s = self.conv_conv
        y = F2.conv_bn_act(x, x_up, s[0], s[1], s[2].negative_slope)
        mask = self._mask_queue.pop(0) if self._mask_queue else None
        y = F2.dropout(y, s[3].p, s[3].training, mask)
        return F2.conv_bn_act(y, None, s[4], s[5], s[6].negative_slope)


class DownBlock(nn.Module):
    """Downsampling followed by ConvBlock"""

    def __init__(self, in_channels, out_channels, dropout_p):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), ConvBlock(in_channels, out_channels, dropout_p))

    def forward(self, x):
        return self.maxpool_conv[1](F2.max_pool2d(x))


class UpBlock(nn.Module):
    """Upsampling followed by ConvBlock"""

    def __init__(self, in_channels1, in_channels2, out_channels, dropout_p, bilinear=True):
        super().__init__()
        if not bilinear:
            raise NotImplementedError("icl_b200 UpBlock implements bilinear=True, the only configuration the reference's Decoder "
                                      "constructs (unet_icl.py:168-175 never passes `bilinear`)")
        self.bilinear = bilinear
        self.conv1x1 = nn.Conv2d(in_channels1, in_channels2, kernel_size=1)
        self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.conv = ConvBlock(in_channels2 * 2, out_channels, dropout_p)

    def forward(self, x1, x2):
        x1 = F2.upsample2x_ac(F2.conv1x1(x1, self.conv1x1))
        return self.conv(x2, x1)


class Encoder(nn.Module):
    def __init__(self, params):
        super().__init__()
        self.params = params
        self.in_chns, self.ft_chns, self.n_class = params["in_chns"], params["feature_chns"], params["class_num"]
        self.bilinear, self.dropout = params["bilinear"], params["dropout"]
        assert len(self.ft_chns) == 5
        self.in_conv = ConvBlock(self.in_chns, self.ft_chns[0], self.dropout[0])
        self.down1 = DownBlock(self.ft_chns[0], self.ft_chns[1], self.dropout[1])
        self.down2 = DownBlock(self.ft_chns[1], self.ft_chns[2], self.dropout[2])
        self.down3 = DownBlock(self.ft_chns[2], self.ft_chns[3], self.dropout[3])
        self.down4 = DownBlock(self.ft_chns[3], self.ft_chns[4], self.dropout[4])

    def forward(self, x):
        x0 = self.in_conv(x)
        x1 = self.down1(x0)
        x2 = self.down2(x1)
        x3 = self.down3(x2)
        x4 = self.down4(x3)
        return [x0, x1, x2, x3, x4]


class Decoder(nn.Module):
    def __init__(self, params, return_feats=False):
        super().__init__()
        self.params = params
        self.in_chns, self.ft_chns, self.n_class = params["in_chns"], params["feature_chns"], params["class_num"]
        self.bilinear = params["bilinear"]
        self.return_feats = return_feats
        assert len(self.ft_chns) == 5
        f = self.ft_chns
        self.up1 = UpBlock(f[4], f[3], f[3], dropout_p=0.0)
        self.up2 = UpBlock(f[3], f[2], f[2], dropout_p=0.0)
        self.up3 = UpBlock(f[2], f[1], f[1], dropout_p=0.0)
        self.up4 = UpBlock(f[1], f[0], f[0], dropout_p=0.0)
        self.out_conv = nn.Conv2d(f[0], self.n_class, kernel_size=3, padding=1)

    def forward(self, feature):
        x0, x1, x2, x3, x4 = feature
        x_1 = self.up1(x4, x3)
        x_2 = self.up2(x_1, x2)
        x_3 = self.up3(x_2, x1)
        x = self.up4(x_3, x0)
        output = F2.conv2d_3x3(x, self.out_conv)
        return (output, [x_1, x_2, x_3]) if self.return_feats else output


class UNet(nn.Module):
    def __init__(self, in_chns, class_num):
        super().__init__()
        params = {"in_chns": in_chns, "feature_chns": [16, 32, 64, 128, 256], "dropout": [0.05, 0.1, 0.2, 0.3, 0.5],
                  "class_num": class_num, "bilinear": False, "acti_func": "relu"}
        self.encoder = Encoder(params)
        self.decoder = Decoder(params)

    def forward(self, x):
        return self.decoder(self.encoder(x))
