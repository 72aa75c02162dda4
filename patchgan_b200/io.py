"""Input pipeline with the reference's dataset conventions (/root/reference/patchgan/io.py:10-58: `<id>.jpg` images and
`<id>.png` label maps, `labels`, `size`, `augmentation`), re-designed so that the per-sample arithmetic runs on the GPU.

The reference does everything in `__getitem__` on CPU workers and hands the trainer float tensors: 4 * (3 + L) bytes per
pixel cross PCIe, and at the 6 k img/s of one B200 the four DataLoader workers of patchgan_train cannot keep up.  Here

  * `COCOStuffDataset.__getitem__` only DECODES: it returns the raw uint8 image (3, H, W), the raw uint8 label map (H, W)
    and the flip decision (drawn with the same probabilities as RandomHorizontalFlip(0.25) / RandomVerticalFlip(0.25));
  * `prepare_batch` uploads those bytes and runs `pg_prep_batch_u8`: /255, the uint8 `+1` label shift, the bilinear
    Resize of the stacked tensor, flips and the per-label masks, with ATen's arithmetic order -- masks are bit-exact, the
    image agrees to 1 ulp (oracle/io_oracle.py, tests/golden/io.npz from the live reference);
  * `DeviceBatches` wraps a DataLoader of such samples and yields (x, y) CUDA float batches, which `Trainer.train`
    consumes like any other loader.

`Trainer.batch` / `submit` also take a raw uint8 batch directly (x: (B,3,H,W) uint8, y: (B,H,W) uint8 label maps) once
`trainer.labels` is set: 4 bytes per pixel of upload instead of 4 * (3 + L) * 4."""
import ctypes
import glob
import os

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from . import _lib as L
from .engine import _stream, require_cuda


def prepare_batch(img_u8, lab_u8, labels, size, flips=None, out=None):
    """img_u8: CUDA uint8 (B,3,H,W); lab_u8: CUDA uint8 (B,H,W); labels: label ids (sorted like io.py:17); size: (S_h, S_w);
    flips: CUDA uint8 (B,) or None.  Returns (x float (B,3,S_h,S_w), y float (B,L,S_h,S_w)); `out` = (x, y) to fill."""
    require_cuda(img_u8, 'img_u8')
    require_cuda(lab_u8, 'lab_u8')
    if img_u8.dtype != torch.uint8 or lab_u8.dtype != torch.uint8:
        raise TypeError('prepare_batch takes raw uint8 images and label maps')
    B, C, H, W = img_u8.shape
    if C != 3 or tuple(lab_u8.shape) != (B, H, W):
        raise RuntimeError(f'expected (B,3,H,W) and (B,H,W), got {tuple(img_u8.shape)} and {tuple(lab_u8.shape)}')
    ids = [int(v) for v in np.sort(np.asarray(labels))]
    Ho, Wo = int(size[0]), int(size[1])
    dev = img_u8.device
    if out is None:
        out = (torch.empty((B, 3, Ho, Wo), device=dev, dtype=torch.float32),
               torch.empty((B, len(ids), Ho, Wo), device=dev, dtype=torch.float32))
    arr = (ctypes.c_int32 * len(ids))(*ids)
    L.call('pg_prep_batch_u8', img_u8.contiguous().data_ptr(), lab_u8.contiguous().data_ptr(), arr, len(ids), B, H, W, Ho, Wo,
           flips.contiguous().data_ptr() if flips is not None else None, out[0].data_ptr(), out[1].data_ptr(), _stream())
    return out


class COCOStuffDataset(Dataset):
    """Same constructor as the reference.  A sample is (image uint8 (3,H,W), label map uint8 (H,W), flip code): the
    decoding is all that happens on the CPU; see prepare_batch / DeviceBatches for the rest."""
    augmentation = None

    def __init__(self, imgfolder, maskfolder, labels=[1], size=256, augmentation='resize'):
        self.images = np.asarray(sorted(glob.glob(os.path.join(imgfolder, "*.jpg"))))
        self.masks = np.asarray(sorted(glob.glob(os.path.join(maskfolder, "*.png"))))
        self.size = size
        self.labels = np.sort(labels)
        stem = lambda f: int(os.path.splitext(os.path.basename(f))[0])
        if [stem(f) for f in self.images] != [stem(f) for f in self.masks]:
            raise AssertionError("Image IDs and Mask IDs do not match!")
        # io.py:22-30: 'randomcrop' resizes, 'randomcrop+flip' resizes and flips; anything else leaves the sample alone
        self.augmentation = augmentation if augmentation in ('randomcrop', 'randomcrop+flip') else None
        print(f"Loaded {len(self)} images")

    def __len__(self):
        return len(self.images)

    def __getitem__(self, index):
        from torchvision.io import ImageReadMode, read_image
        img = read_image(self.images[index], ImageReadMode.RGB)
        lab = read_image(self.masks[index], ImageReadMode.GRAY)[0]
        flip = 0
        if self.augmentation == 'randomcrop+flip':
            flip = int(torch.rand(1) < 0.25) | (int(torch.rand(1) < 0.25) << 1)
        return img, lab, flip

    def out_size(self, h, w):
        return (self.size, self.size) if self.augmentation is not None else (h, w)


class DeviceBatches:
    """Iterable of (x, y) CUDA float batches over a dataset of raw samples: a DataLoader (workers decode, batches stay
    lists because the raw sizes differ) + one upload and one kernel launch per group of equally sized samples."""

    def __init__(self, dataset, batch_size=16, shuffle=True, num_workers=4, device='cuda', **loader_kwargs):
        self.base = dataset.dataset if hasattr(dataset, 'dataset') else dataset          # (random_split wraps it)
        self.device = torch.device(device)
        if num_workers > 0:
            loader_kwargs.setdefault('persistent_workers', True)
        self.loader = DataLoader(dataset, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers,
                                 collate_fn=list, **loader_kwargs)

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        for samples in self.loader:
            yield self.to_device(samples)

    def to_device(self, samples):
        B = len(samples)
        h0, w0 = samples[0][0].shape[1:]
        Ho, Wo = self.base.out_size(h0, w0)
        nl = len(self.base.labels)
        x = torch.empty((B, 3, Ho, Wo), device=self.device, dtype=torch.float32)
        y = torch.empty((B, nl, Ho, Wo), device=self.device, dtype=torch.float32)
        flips = torch.tensor([s[2] for s in samples], dtype=torch.uint8)
        i = 0
        while i < B:                      # runs of equal raw size go up together
            j = i + 1
            while j < B and samples[j][0].shape == samples[i][0].shape:
                j += 1
            img = torch.stack([s[0] for s in samples[i:j]]).pin_memory().to(self.device, non_blocking=True)
            lab = torch.stack([s[1] for s in samples[i:j]]).pin_memory().to(self.device, non_blocking=True)
            if self.base.out_size(*img.shape[2:]) != (Ho, Wo):
                raise RuntimeError('samples of different sizes need augmentation="randomcrop" (a common output size)')
            prepare_batch(img, lab, self.base.labels, (Ho, Wo), flips[i:j].to(self.device), out=(x[i:j], y[i:j]))
            i = j
        return x, y
