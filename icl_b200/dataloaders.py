"""Batch assembly for the 3D ICL loops with the volumes RESIDENT ON THE DEVICE — SURVEY.md §8f item 4.

Mirrors the pieces of the reference's dataloaders/brats2019.py that `train_inherent_consistent_unet_3D_BraTS.py:66-83` uses:
`TwoStreamBatchSampler` (:191-218 with iterate_once / iterate_eternally / grouper :221-236), `RandomRotFlip` (:133-148),
`RandomCrop` (:80-130), `CenterCrop` (:48-77), `ToTensor` (:176-188).  The reference runs them on numpy arrays in 4 DataLoader
worker processes and ships every 4 x 96^3 batch host -> device (14 MB per step; at 23 ms per step that is a 0.6 GB/s stream
per GPU plus the worker CPU time, and 8 GPUs share one host).  Here the transforms take torch tensors on ANY device and are
pure index arithmetic (pad / rot90 / flip / slice), so a dataset that fits HBM (BraTS2019: 250 training volumes,
data/BraTS2019/train.txt; brain-bbox crops of a 240x240x155 grid, i.e. at most 36 MB each as fp32 — under 9 GB of the 180 GB)
stays on the GPU and a step's batch is cut out of it without touching the host.

Random numbers are drawn from numpy's global generator in exactly the reference's order (RotFlip: k = randint(0, 4) then
axis = randint(0, 2); RandomCrop: w1, h1, d1 = randint(0, dim - out)), so with the same `np.random.seed` the outputs are
bit-identical to the reference's transforms (tests/test_dataloaders_cpu.py)."""
import itertools

import numpy as np
import torch
import torch.nn.functional as F


def _as_tensor(a, device=None):
    t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
    return t.to(device) if device is not None else t


def _pad_to(image, label, output_size):
    """Zero-pad both volumes when any label dimension is <= the crop size: (out - dim) // 2 + 3 voxels per side, never negative."""
    if any(label.shape[i] <= output_size[i] for i in range(3)):
        p = [max((output_size[i] - label.shape[i]) // 2 + 3, 0) for i in range(3)]
        pads = (p[2], p[2], p[1], p[1], p[0], p[0])  # F.pad lists the last dimension first
        image = F.pad(image, pads)
        label = F.pad(label, pads)
    return image, label


class RandomRotFlip(object):
    """k quarter turns in the (0, 1) plane, then a flip along axis 0 or 1 (brats2019.py:140-146)."""

    def __call__(self, sample):
        image, label = _as_tensor(sample["image"]), _as_tensor(sample["label"])
        k = int(np.random.randint(0, 4))
        axis = int(np.random.randint(0, 2))
        image = torch.flip(torch.rot90(image, k, dims=(0, 1)), dims=(axis,))
        label = torch.flip(torch.rot90(label, k, dims=(0, 1)), dims=(axis,))
        return {"image": image, "label": label}


class RandomCrop(object):
    """Random output_size window after the optional zero padding (brats2019.py:91-130; the signed-distance branch is not on the ICL path)."""

    def __init__(self, output_size, with_sdf=False):
        if with_sdf:
            raise NotImplementedError("icl_b200 RandomCrop: with_sdf is not used by the ICL loops")
        self.output_size = tuple(int(v) for v in output_size)

    def __call__(self, sample):
        image, label = _pad_to(_as_tensor(sample["image"]), _as_tensor(sample["label"]), self.output_size)
        w, h, d = image.shape
        o = self.output_size
        w1 = int(np.random.randint(0, w - o[0]))
        h1 = int(np.random.randint(0, h - o[1]))
        d1 = int(np.random.randint(0, d - o[2]))
        return {"image": image[w1:w1 + o[0], h1:h1 + o[1], d1:d1 + o[2]], "label": label[w1:w1 + o[0], h1:h1 + o[1], d1:d1 + o[2]]}


class CenterCrop(object):
    """brats2019.py:48-77."""

    def __init__(self, output_size):
        self.output_size = tuple(int(v) for v in output_size)

    def __call__(self, sample):
        image, label = _pad_to(_as_tensor(sample["image"]), _as_tensor(sample["label"]), self.output_size)
        w, h, d = image.shape
        o = self.output_size
        w1, h1, d1 = int(round((w - o[0]) / 2.0)), int(round((h - o[1]) / 2.0)), int(round((d - o[2]) / 2.0))
        return {"image": image[w1:w1 + o[0], h1:h1 + o[1], d1:d1 + o[2]], "label": label[w1:w1 + o[0], h1:h1 + o[1], d1:d1 + o[2]]}


class ToTensor(object):
    """image -> float32 [1, w, h, d], label -> int64 [w, h, d] (brats2019.py:179-188); both stay on the sample's device."""

    def __call__(self, sample):
        image, label = _as_tensor(sample["image"]), _as_tensor(sample["label"])
        return {"image": image.reshape((1,) + tuple(image.shape)).to(torch.float32).contiguous(), "label": label.long().contiguous()}


class Compose(object):
    """torchvision.transforms.Compose (train_inherent_consistent_unet_3D_BraTS.py:69-73) without the torchvision import."""

    def __init__(self, transforms):
        self.transforms = list(transforms)

    def __call__(self, sample):
        for t in self.transforms:
            sample = t(sample)
        return sample


class TwoStreamBatchSampler(torch.utils.data.Sampler):
    """Batches of (batch_size - secondary_batch_size) primary (labeled) + secondary_batch_size secondary (unlabeled) indices.
    One epoch = one pass over a permutation of the primary indices; the secondary indices are re-permuted for ever
    (brats2019.py:191-236).  np.random.permutation is called in the reference's order: primary first, then one secondary
    permutation each time the previous one is used up."""

    def __init__(self, primary_indices, secondary_indices, batch_size, secondary_batch_size):
        self.primary_indices = primary_indices
        self.secondary_indices = secondary_indices
        self.secondary_batch_size = secondary_batch_size
        self.primary_batch_size = batch_size - secondary_batch_size
        assert len(self.primary_indices) >= self.primary_batch_size > 0
        assert len(self.secondary_indices) >= self.secondary_batch_size > 0

    def _secondary_stream(self):
        while True:
            for i in np.random.permutation(self.secondary_indices):
                yield i

    def __iter__(self):
        primary = iter(np.random.permutation(self.primary_indices))  # drawn now, like iterate_once at :206
        secondary = self._secondary_stream()                         # first permutation drawn when first consumed

        def batches():
            while True:
                p = tuple(itertools.islice(primary, self.primary_batch_size))
                if len(p) < self.primary_batch_size:  # incomplete last primary group is dropped
                    return
                s = tuple(itertools.islice(secondary, self.secondary_batch_size))
                yield p + s
        return batches()

    def __len__(self):
        return len(self.primary_indices) // self.primary_batch_size


class DeviceVolumeSet(object):
    """The role of `BraTS2019(..., transform=Compose([RandomRotFlip(), RandomCrop(patch), ToTensor()]))` + DataLoader(batch_sampler=...)
    (train_inherent_consistent_unet_3D_BraTS.py:66-83) with every volume kept on `device`.

    volumes: sequence of (image [w,h,d] float, label [w,h,d] integer) arrays / tensors.  `batch(indices)` applies the transform per
    sample on the device and stacks: {'image': [B,1,pw,ph,pd] float32, 'label': [B,pw,ph,pd] int64} — the tensors the loop
    indexes at :98-99.  Iterating yields one such batch per index group of the sampler."""

    def __init__(self, volumes, transform=None, device="cuda", batch_sampler=None):
        self.device = torch.device(device)
        self.items = [(_as_tensor(im, self.device).to(torch.float32), _as_tensor(lb, self.device).to(torch.uint8)) for im, lb in volumes]
        self.transform = transform
        self.batch_sampler = batch_sampler

    def __len__(self):
        return len(self.items) if self.batch_sampler is None else len(self.batch_sampler)

    def __getitem__(self, idx):
        image, label = self.items[int(idx)]
        sample = {"image": image, "label": label}
        return self.transform(sample) if self.transform else sample

    def batch(self, indices):
        samples = [self[i] for i in indices]
        return {"image": torch.stack([s["image"] for s in samples]), "label": torch.stack([s["label"] for s in samples])}

    def __iter__(self):
        if self.batch_sampler is None:
            raise RuntimeError("DeviceVolumeSet: iteration needs a batch_sampler")
        for indices in self.batch_sampler:
            yield self.batch(indices)
