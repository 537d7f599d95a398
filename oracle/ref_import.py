"""Import the UNMODIFIED reference (read-only, /root/reference/code) for oracle pinning.

Only usable in the build container; the GPU box has no /root/reference.  Follows the recipe
verified in SURVEY.md Appendix C: stub modules for monai (and, for the Swin files, timm and turtle) first on sys.path,
then the reference's ``code`` directory.  Never imports ``networks.net_factory*`` (argv parsing and
missing modules at import, net_factory_3d.py:3-37).
"""
import importlib
import os
import sys
import types

# /root/reference exists in the build container only; baseline/_ref (git-ignored, shipped by gpurun) is a verbatim copy of the
# reference's code/{networks,utils} + the two eval scripts, made by __graft_entry__.build() so that the reference arm of bench.py
# can run the UNMODIFIED reference on the GPU box too.
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = [os.environ.get("ICL_REFERENCE_ROOT", "/root/reference"), os.path.join(_REPO, "baseline", "_ref")]
REF_ROOT = next((r for r in _CANDIDATES if os.path.isdir(os.path.join(r, "code", "networks"))), _CANDIDATES[0])
REF_CODE = os.path.join(REF_ROOT, "code")
STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_CODE, "networks"))


def _ensure_path():
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_CODE)
    for p in (REF_CODE, STUBS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = [STUBS, REF_CODE]
    # The product package also has sub-packages called ``networks``/``utils`` *inside*
    # icl_b200; the reference uses top-level ``networks``/``utils``.  No clash.


def load():
    """Returns a namespace with the reference classes on the hot path."""
    _ensure_path()
    ns = types.SimpleNamespace()
    ns.unet_3D = importlib.import_module("networks.unet_3D").unet_3D
    m = importlib.import_module("networks.unet_3D_icl")
    ns.unet_3D_icl = m.unet_3D_icl
    ns.InherentConsistent = m.InherentConsistent
    ns.Class_Decoder = m.Class_Decoder
    ns.Query_Attention = m.Query_Attention
    ns.SeparableConv3d = m.SeparableConv3d
    ns.MLP = m.MLP
    m2 = importlib.import_module("networks.unet_icl")   # 2D path (config 1)
    ns.UNet_icl = m2.UNet_icl
    ns.Encoder2d, ns.Decoder2d = m2.Encoder, m2.Decoder
    ns.InherentConsistent2d = m2.InherentConsistent
    ns.UNet = importlib.import_module("networks.unet").UNet
    m3 = importlib.import_module("networks.vision_transformer")   # Swin path (config 4); timm / turtle come from oracle/stubs
    ns.SwinUnet = m3.SwinUnet
    ns.InherentConsistentTokens = m3.InherentConsistent
    m4 = importlib.import_module("networks.swinunet_icl")
    ns.SwinTransformerSys, ns.SwinTransformerBlock, ns.WindowAttention = m4.SwinTransformerSys, m4.SwinTransformerBlock, m4.WindowAttention
    ns.PatchMerging, ns.PatchExpand, ns.FinalPatchExpand_X4, ns.PatchEmbed = m4.PatchMerging, m4.PatchExpand, m4.FinalPatchExpand_X4, m4.PatchEmbed
    ns.net_utils = importlib.import_module("networks.utils")
    ns.losses = importlib.import_module("utils.losses")
    return ns


def load_test_single_case():
    """Reference sliding-window driver test_3D_BraTS.test_single_case (test_3D_BraTS.py:79-142).

    The module imports h5py/medpy/SimpleITK/tqdm/net_factory_3d at top level; none is used by
    ``test_single_case`` itself, so empty stand-in modules are registered before import.
    """
    _ensure_path()
    for name in ("h5py", "nibabel", "SimpleITK", "medpy", "medpy.metric", "skimage", "skimage.measure",
                 "networks.net_factory_3d"):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            sys.modules[name] = mod
    sys.modules["medpy"].metric = sys.modules["medpy.metric"]
    sys.modules["skimage.measure"].label = None
    sys.modules["networks.net_factory_3d"].net_factory_3d = None
    argv, sys.argv = sys.argv, sys.argv[:1]  # the script parses argv at import (test_3D_BraTS.py:21-33)
    try:
        mod = importlib.import_module("test_3D_BraTS")
    finally:
        sys.argv = argv
    return mod.test_single_case
