"""CPU: the C-ABI library builds, loads and exports every symbol include/icl_b200.h declares; the product path
refuses to run without CUDA (no CPU fallback); module trees match the reference's state_dict keys."""
import json
import os
import subprocess

import pytest
import torch

from helpers import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from icl_b200 import build
    return build.build()


def test_exports_match_header(lib_path):
    from icl_b200 import _lib
    protos = _lib.header_prototypes()
    assert len(protos) >= 40
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = set(protos) - exported
    assert not missing, missing
    extra = {s for s in exported if s.startswith("icl_")} - set(protos)
    assert not extra, "exported but undeclared: %s" % extra


def test_library_loads_without_gpu(lib_path):
    from icl_b200 import _lib
    l = _lib.lib()
    assert l.icl_version() >= 100
    assert l.icl_umma_ntile(16) == 16 and l.icl_umma_ntile(48) == 48 and l.icl_umma_ntile(256) == 128
    assert l.icl_umma_ntile(384) == 128 and l.icl_umma_ntile(192) == 96 and l.icl_umma_ntile(20) == 0
    assert l.icl_sgd_chunk() > 0


def test_sass_has_blackwell_tensor_path(lib_path):
    sass = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UBLKCP"):
        assert mnemonic in sass, mnemonic


def test_no_cpu_fallback():
    from icl_b200.networks.unet_3D import unet_3D
    from icl_b200.utils import losses
    net = unet_3D(feature_scale=16, n_classes=2, in_channels=1)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 1, 16, 16, 16))
    with pytest.raises(RuntimeError, match="CUDA"):
        losses.CrossEntropyLoss()(torch.zeros(1, 2, 4, 4, 4), torch.zeros(1, 4, 4, 4, dtype=torch.long))


def test_state_dict_keys_match_reference_fixture(monkeypatch):
    monkeypatch.setattr(torch.nn.init, "kaiming_normal_", lambda t, **k: t)
    monkeypatch.setattr(torch.nn.init, "kaiming_uniform_", lambda t, **k: t)
    from icl_b200.networks.unet_3D import unet_3D
    from icl_b200.networks.unet_3D_icl import unet_3D_icl
    keys = json.load(open(os.path.join(GOLDEN, "state_keys.json")))
    m = unet_3D_icl(feature_scale=4, n_classes=2, in_channels=1)
    assert [[k, list(v.shape)] for k, v in m.state_dict().items()] == keys["unet_3D_icl_k2"]
    assert [k for k, _ in m.named_parameters()] == keys["unet_3D_icl_k2_params"]
    m2 = unet_3D(feature_scale=4, n_classes=2, in_channels=1)
    assert [[k, list(v.shape)] for k, v in m2.state_dict().items()] == keys["unet_3D_k2"]
    # ICL checkpoints are saved without sspa/uscl keys and loaded strictly into unet_3D (train_..._BraTS.py:158-162, test_3D_BraTS.py:147-150)
    ck = {k: v for k, v in m.state_dict().items() if "sspa" not in k and "uscl" not in k}
    m2.load_state_dict(ck, strict=True)


@pytest.mark.reference
def test_state_dict_interchange_with_live_reference(monkeypatch):
    from oracle import ref_import
    ns = ref_import.load()
    from icl_b200.networks.unet_3D import unet_3D
    ref = ns.unet_3D(feature_scale=4, n_classes=3, in_channels=2)
    mine = unet_3D(feature_scale=4, n_classes=3, in_channels=2)
    mine.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(mine.state_dict(), strict=True)


def test_2d_and_swin_mirrors_refuse_cpu_tensors():
    """No CPU fallback on the 2D / Swin paths either: the first kernel-backed op raises."""
    from icl_b200.networks.unet_icl import UNet_icl
    from icl_b200.networks.vision_transformer import SwinUnet, swin_tiny_lite_config
    with pytest.raises(RuntimeError, match="CUDA"):
        UNet_icl(1, 4)(torch.zeros(1, 1, 32, 32), torch.zeros(1, 1, 32, 32))
    with pytest.raises(RuntimeError, match="CUDA"):
        SwinUnet(swin_tiny_lite_config(), img_size=224, num_classes=4)(torch.zeros(1, 1, 224, 224), inference=True)
