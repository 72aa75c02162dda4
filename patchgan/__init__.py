"""Drop-in alias: ``import patchgan`` (the reference's package name, /root/reference/patchgan/__init__.py:1-8) resolves
to the B200 implementation, sub-modules included (``patchgan.unet``, ``patchgan.trainer``, ``patchgan.losses`` ...), so
code written against the reference runs unchanged.  Put this repository BEFORE any installed reference on sys.path."""
import importlib
import sys

from patchgan_b200 import Discriminator, Trainer, UNet, __version__  # noqa: F401

__all__ = ['UNet', 'Discriminator', 'Trainer', '__version__']

for _name in ('unet', 'disc', 'trainer', 'losses', 'transfer', 'io', 'train', 'infer', 'version'):
    sys.modules[f'{__name__}.{_name}'] = importlib.import_module(f'patchgan_b200.{_name}')
    globals()[_name] = sys.modules[f'{__name__}.{_name}']
