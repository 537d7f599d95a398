"""icl_b200 — B200-native (sm_100a) implementation of the ICL training / inference hot path.

Mirrors the reference's Python API for that path (zhuye98/ICL, code/networks + code/utils/losses.py):
    from icl_b200.networks.unet_3D import unet_3D
    from icl_b200.networks.unet_3D_icl import unet_3D_icl
    from icl_b200.utils import losses
Every op underneath is a hand-written CUDA kernel reached through the C-ABI in include/icl_b200.h.
"""
from .precision import get_precision, set_precision  # noqa: F401

__version__ = "0.1.0"
