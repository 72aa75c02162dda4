"""``patchgan_train`` console script (reference: /root/reference/patchgan/train.py:13-127): argparse + YAML ->
Dataset / DataLoader / UNet / Discriminator / Trainer.train().  Host glue only; the work happens in Trainer.batch."""
import argparse

import torch
import yaml
from torch.utils.data import DataLoader, random_split

from . import config as cfg
from .disc import Discriminator
from .trainer import Trainer
from .unet import UNet


def build_models(config, in_channels, out_channels, device):
    mp = cfg.model_params(config)
    generator = UNet(in_channels, out_channels, mp['gen_filts'], use_dropout=mp['use_dropout'],
                     activation=mp['activation'], final_act=mp['final_activation']).to(device)
    discriminator = Discriminator(in_channels + out_channels, mp['disc_filts'], norm=mp['disc_norm'],
                                  n_layers=mp['n_disc_layers']).to(device)
    return generator, discriminator


def patchgan_train():
    parser = argparse.ArgumentParser(prog='PatchGAN', description='Train the PatchGAN architecture')
    parser.add_argument('-c', '--config_file', required=True, type=str, help='Location of the config YAML file')
    parser.add_argument('-b', '--batch_size', default=16, type=int, help='Number of images per batch')
    parser.add_argument('--dataloader_workers', default=4, type=int,
                        help='Number of workers to use with dataloader (set to 0 to disable multithreading)')
    parser.add_argument('-n', '--n_epochs', required=True, type=int, help='Number of epochs to train the model')
    parser.add_argument('-d', '--device', default='auto', help='Device to use to train the model (CUDA=GPU)')
    parser.add_argument('--summary', default=True, action='store_true', help="Print summary of the models")
    args = parser.parse_args()

    device = 'cuda' if args.device == 'auto' else args.device
    if device != 'cuda' or not torch.cuda.is_available():
        raise SystemExit('patchgan_b200 trains on a CUDA (sm_100a) device only')

    with open(args.config_file, 'r') as infile:
        config = yaml.safe_load(infile)
    ds = config['dataset']
    Dataset, in_channels, out_channels, ds_kwargs = cfg.dataset_class(ds)
    size, augmentation = ds.get('size', 256), ds.get('augmentation', 'randomcrop')
    train_paths, val_paths, split = cfg.data_paths(config)
    if split is None:
        train_set = Dataset(train_paths['images'], train_paths['masks'], size=size, augmentation=augmentation, **ds_kwargs)
        val_set = Dataset(val_paths['images'], val_paths['masks'], size=size, augmentation=augmentation, **ds_kwargs)
    else:
        full = Dataset(train_paths['images'], train_paths['masks'], size=size, augmentation=augmentation, **ds_kwargs)
        train_set, val_set = random_split(full, split)

    from .io import COCOStuffDataset, DeviceBatches
    if Dataset is COCOStuffDataset:
        # raw uint8 samples from the workers; /255, label shift, resize, flips and masks on the GPU (patchgan_b200/io.py)
        train_data = DeviceBatches(train_set, args.batch_size, True, args.dataloader_workers, device)
        val_data = DeviceBatches(val_set, args.batch_size, True, args.dataloader_workers, device)
    else:
        loader_kwargs = dict(num_workers=args.dataloader_workers, persistent_workers=True) if args.dataloader_workers > 0 else {}
        train_data = DataLoader(train_set, batch_size=args.batch_size, shuffle=True, pin_memory=True, **loader_kwargs)
        val_data = DataLoader(val_set, batch_size=args.batch_size, shuffle=True, pin_memory=True, **loader_kwargs)

    generator, discriminator = build_models(config, in_channels, out_channels, device)
    if args.summary:
        for name, net in (('generator', generator), ('discriminator', discriminator)):
            print(f'{name}: {sum(p.numel() for p in net.parameters()):,} parameters')

    trainer = Trainer(generator, discriminator, savefolder=config.get('checkpoint_path', './checkpoints/'))
    if config.get('load_last_checkpoint', False):
        trainer.load_last_checkpoint()
    elif config.get('transfer_learn', {}).get('generator_checkpoint', None) is not None:
        tl = config['transfer_learn']
        generator.load_transfer_data(torch.load(tl['generator_checkpoint'], map_location=device))
        discriminator.load_transfer_data(torch.load(tl['discriminator_checkpoint'], map_location=device))

    tp = config['train_params']
    trainer.loss_type = tp['loss_type']
    trainer.seg_alpha = tp['seg_alpha']
    trainer.train(train_data, val_data, args.n_epochs, dsc_learning_rate=tp['disc_learning_rate'],
                  gen_learning_rate=tp['gen_learning_rate'], lr_decay=tp.get('decay_rate', None),
                  save_freq=tp.get('save_freq', 10))
