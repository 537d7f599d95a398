"""`net_factory` — drop-in for the reference's networks/net_factory.py:78-89, without its import-time argparse (the reference
module parses sys.argv and builds `config` when imported, :10-75).  `config` / `num_classes` are explicit keyword arguments here;
the defaults are what the reference's defaults produce (swin-tiny "lite" yaml, class_num)."""
from .unet import UNet
from .unet_icl import UNet_icl
from .vision_transformer import SwinUnet, swin_tiny_lite_config


def net_factory(net_type="unet", in_chns=1, class_num=3, config=None):
    if net_type == "unet":
        return UNet(in_chns=in_chns, class_num=class_num).cuda()
    if net_type == "icl_unet":
        return UNet_icl(in_chns=in_chns, class_num=class_num).cuda()
    if net_type == "icl_swinunet":
        return SwinUnet(config or swin_tiny_lite_config(), img_size=[224, 224], num_classes=class_num).cuda()
    if net_type == "swinunet":
        raise NotImplementedError("icl_b200 builds the ICL path; the plain Swin-UNet of networks/vision_transformer_base.py is the "
                                  "`swin_unet` sub-module of 'icl_swinunet' (SwinUnet(...)(x, inference=True))")
    return None
