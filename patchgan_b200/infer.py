"""Tiled inference (reference: /root/reference/patchgan/infer.py:14-174): cut an image into overlapping crops, run the
generator, overlap-average / threshold / argmax the crops back into a mask.  ``n_crop`` and ``build_mask`` run as
device kernels (pg_ncrop / pg_build_mask) and reproduce the reference bit for bit -- including its
``j * ncropsy + i`` crop index (infer.py:32,57), which differs from the row-major index when the image is not square."""
import argparse
import ctypes
import os

import numpy as np
import torch
import tqdm
import yaml

from . import _lib as L
from . import config as cfg
from .engine import _stream, require_cuda


def _grid(height, width, size, overlap):
    eff = int(overlap * size)
    return eff, int(np.ceil(height / eff)), int(np.ceil(width / eff))


def n_crop(image, size, overlap):
    """image: CUDA float tensor (C, H, W) -> crops (ncropsx*ncropsy, C, size, size) on the same device."""
    require_cuda(image, 'image')
    c, height, width = image.shape
    eff, ncy, ncx = _grid(height, width, size, overlap)
    if (ncy - 1) * ncy + ncx - 1 >= ncx * ncy:
        raise IndexError(f'index {(ncy - 1) * ncy + ncx - 1} is out of bounds for dimension 0 with size {ncx * ncy}')
    img = image.contiguous().float()
    crops = torch.empty((ncx * ncy, c, size, size), device=image.device, dtype=torch.float32)
    L.call('pg_ncrop', img.data_ptr(), crops.data_ptr(), c, height, width, size, eff, ncy, ncx, _stream())
    return crops


def build_mask(masks, crop_size, image_size, threshold, overlap):
    """masks: CUDA float tensor (n, C, crop, crop) from the generator -> numpy mask like the reference
    (argmax over channels, int64, if C > 1; else the averaged / thresholded single channel, float64)."""
    require_cuda(masks, 'masks')
    n, c, height, width = masks.shape
    ih, iw = image_size
    eff, ncy, ncx = _grid(ih, iw, crop_size, overlap)
    m = masks.contiguous().float()
    out = torch.empty((c, ih, iw), device=masks.device, dtype=torch.float32)
    arg = torch.empty((ih, iw), device=masks.device, dtype=torch.int32) if c > 1 else None
    L.call('pg_build_mask', m.data_ptr(), out.data_ptr(), arg.data_ptr() if arg is not None else None, c, ih, iw,
           crop_size, eff, ncy, ncx, float(threshold), _stream())
    if c > 1:
        return arg.cpu().numpy().astype(np.int64)
    return out[0].cpu().numpy().astype(np.float64)


def patchgan_infer():
    from .train import build_models
    parser = argparse.ArgumentParser(prog='PatchGAN', description='Run inference with the PatchGAN generator')
    parser.add_argument('-c', '--config_file', required=True, type=str, help='Location of the config YAML file')
    parser.add_argument('--dataloader_workers', default=4, type=int)
    parser.add_argument('-d', '--device', default='auto', help='Device to use (CUDA=GPU)')
    parser.add_argument('--summary', default=True, action='store_true', help="Print summary of the models")
    args = parser.parse_args()
    device = 'cuda' if args.device == 'auto' else args.device
    if device != 'cuda' or not torch.cuda.is_available():
        raise SystemExit('patchgan_b200 runs on a CUDA (sm_100a) device only')
    print(f"Running with {device}")

    with open(args.config_file, 'r') as infile:
        config = yaml.safe_load(infile)
    ds = config['dataset']
    size = ds.get('size', 256)
    Dataset, in_channels, out_channels, ds_kwargs = cfg.dataset_class(ds)
    for method in ('get_filename', 'save_mask'):
        assert callable(getattr(Dataset, method, None)), \
            f"Dataset class {Dataset.__name__} must have the {method} method"
    datagen = Dataset(ds['dataset_path'], **ds_kwargs)

    generator, discriminator = build_models(config, in_channels, out_channels, device)
    paths = config['checkpoint_paths']
    infer_params = config.get('infer_params', {})
    output_path = infer_params.get('output_path', 'predictions/')
    if not os.path.exists(output_path):
        os.makedirs(output_path)
        print(f"Created folder {output_path}")
    generator.eval()
    discriminator.eval()
    generator.load_state_dict(torch.load(paths['generator'], map_location=device))
    discriminator.load_state_dict(torch.load(paths['discriminator'], map_location=device))
    threshold, overlap = infer_params.get('threshold', 0), infer_params.get('overlap', 0.9)

    for i, data in enumerate(tqdm.tqdm(datagen, desc='Predicting', dynamic_ncols=True, ascii=True)):
        image = torch.as_tensor(data, dtype=torch.float32).to(device)
        out_fname, _ = os.path.splitext(datagen.get_filename(i))
        with torch.no_grad():
            masks = generator(n_crop(image, size, overlap))
        mask = build_mask(masks, size, tuple(image.shape[1:]), threshold, overlap)
        Dataset.save_mask(mask, output_path, out_fname)
