"""Dataset with the reference's semantics (/root/reference/patchgan/io.py:10-58): jpg image / png label-mask pairs,
image scaled to [0,1], labels shifted by +1, resize (+ random flips), one binary mask per requested label.
CPU data loading is outside the accelerated hot path; this class exists so that
``patchgan_train`` runs unchanged."""
import glob
import os

import numpy as np
import torch
from torch.utils.data import Dataset


class COCOStuffDataset(Dataset):
    augmentation = None

    def __init__(self, imgfolder, maskfolder, labels=[1], size=256, augmentation='resize'):
        from torchvision import transforms
        self.images = np.asarray(sorted(glob.glob(os.path.join(imgfolder, "*.jpg"))))
        self.masks = np.asarray(sorted(glob.glob(os.path.join(maskfolder, "*.png"))))
        self.size = size
        self.labels = np.sort(labels)
        ids = [[int(os.path.splitext(os.path.basename(f))[0]) for f in group] for group in (self.images, self.masks)]
        assert ids[0] == ids[1], "Image IDs and Mask IDs do not match!"
        resize = transforms.Resize(size=(size, size), antialias=None)
        if augmentation == 'randomcrop':
            self.augmentation = resize
        elif augmentation == 'randomcrop+flip':
            self.augmentation = transforms.Compose([resize, transforms.RandomHorizontalFlip(0.25),
                                                    transforms.RandomVerticalFlip(0.25)])
        print(f"Loaded {len(self)} images")

    def __len__(self):
        return len(self.images)

    def __getitem__(self, index):
        from torchvision.io import ImageReadMode, read_image
        img = read_image(self.images[index], ImageReadMode.RGB) / 255.
        labels = read_image(self.masks[index], ImageReadMode.GRAY) + 1
        stacked = torch.cat((img, labels), dim=0)
        if self.augmentation is not None:
            stacked = self.augmentation(stacked)
        img, labels = stacked[:3, :], stacked[3, :]
        mask = torch.zeros((len(self.labels), labels.shape[0], labels.shape[1]))
        for i, label in enumerate(self.labels):
            mask[i, labels == label] = 1
        return img, mask
