"""GPU: the on-device batch assembly (icl_b200/dataloaders.py, SURVEY section 8f item 4: TwoStreamBatchSampler + RandomRotFlip /
RandomCrop / CenterCrop / ToTensor of dataloaders/brats2019.py:48-236 on volumes resident in HBM).  The CPU path of the same classes
is pinned bit-identical to the live reference by tests/test_dataloaders_cpu.py; here the CUDA path must reproduce the CPU path
bit for bit under the same numpy seed, and a batch must feed the training step directly."""
import numpy as np
import pytest
import torch

from icl_b200 import dataloaders as D

pytestmark = pytest.mark.gpu


def _volume(shape, seed):
    rng = np.random.RandomState(seed)
    return rng.randn(*shape).astype(np.float32), rng.randint(0, 3, size=shape).astype(np.uint8)


@pytest.mark.parametrize("shape", [(130, 120, 110), (96, 140, 100), (60, 70, 155)])
def test_device_transforms_match_cpu_path(shape):
    image, label = _volume(shape, 3)
    patch = (96, 96, 96)
    tf = D.Compose([D.RandomRotFlip(), D.RandomCrop(patch), D.ToTensor()])
    img_d, lab_d = torch.from_numpy(image).cuda(), torch.from_numpy(label).cuda()
    for seed in (0, 1, 2):
        np.random.seed(seed)
        want = tf({"image": image, "label": label})
        np.random.seed(seed)
        got = tf({"image": img_d, "label": lab_d})
        assert got["image"].is_cuda and got["label"].is_cuda
        assert got["image"].dtype == torch.float32 and got["label"].dtype == torch.int64
        assert torch.equal(got["image"].cpu(), want["image"]) and torch.equal(got["label"].cpu(), want["label"])
    want_c = D.CenterCrop(patch)({"image": image, "label": label})
    got_c = D.CenterCrop(patch)({"image": img_d, "label": lab_d})
    assert torch.equal(torch.as_tensor(got_c["image"]).cpu(), torch.as_tensor(want_c["image"]))
    assert torch.equal(torch.as_tensor(got_c["label"]).cpu().long(), torch.as_tensor(want_c["label"]).long())


def test_device_volume_set_feeds_the_model():
    """Volumes live in HBM; one epoch of TwoStream batches (labeled first, then unlabeled: train_..._BraTS.py:77-80) equals the CPU
    path under the same seed, and a batch goes straight into unet_3D (no host round trip)."""
    from icl_b200.networks.unet_3D import unet_3D
    from icl_b200.utils import synth
    vols = [_volume((100 + 3 * i, 110, 98 + i), i) for i in range(6)]
    tf = D.Compose([D.RandomRotFlip(), D.RandomCrop((96, 96, 96)), D.ToTensor()])
    mk = lambda dev: D.DeviceVolumeSet(vols, transform=tf, device=dev, batch_sampler=D.TwoStreamBatchSampler([0, 1], [2, 3, 4, 5], 4, 2))
    np.random.seed(4)
    cpu = list(mk("cpu"))
    np.random.seed(4)
    gpu = list(mk("cuda"))
    assert len(cpu) == len(gpu) == 1
    b = gpu[0]
    assert b["image"].is_cuda and tuple(b["image"].shape) == (4, 1, 96, 96, 96) and tuple(b["label"].shape) == (4, 96, 96, 96)
    assert torch.equal(b["image"].cpu(), cpu[0]["image"]) and torch.equal(b["label"].cpu(), cpu[0]["label"])
    net = unet_3D(feature_scale=4, n_classes=3, in_channels=1)
    synth.load_synth(net, 5)
    net.cuda().eval()
    with torch.no_grad():
        out = net(b["image"][:1])
    assert tuple(out.shape) == (1, 3, 96, 96, 96) and torch.isfinite(out).all()
