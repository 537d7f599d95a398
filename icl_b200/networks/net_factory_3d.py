"""`net_factory_3d` — drop-in for the reference's networks/net_factory_3d.py:39-68 for the two networks on the ICL hot path
(`unet_3D`, `unet_3D_icl`); the reference module also parses sys.argv at import (:9-37), which this one does not."""
from .unet_3D import unet_3D
from .unet_3D_icl import unet_3D_icl


def net_factory_3d(net_type="unet_3D", in_chns=1, class_num=2):
    if net_type == "unet_3D":
        return unet_3D(n_classes=class_num, in_channels=in_chns).cuda()
    if net_type == "unet_3D_icl":
        return unet_3D_icl(n_classes=class_num, in_channels=in_chns).cuda()
    if net_type in ("swinunetr", "swinunetr_icl", "nnUNet"):
        raise NotImplementedError("icl_b200: %s is outside the scope table (SURVEY.md §8f item 4)" % net_type)
    return None
