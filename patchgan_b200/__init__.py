"""patchgan_b200: B200-native (sm_100a) implementation of the patchGAN training / inference hot path behind the
reference's Python API (UNet, Discriminator, Trainer, losses).  Kernels live in libpatchgan_b200.so (C-ABI in
include/patchgan_b200.h); there is no CPU or library fallback."""
from .unet import UNet
from .disc import Discriminator
from .trainer import Trainer
from .version import __version__

__all__ = [
    'UNet', 'Discriminator', 'Trainer', '__version__'
]
