"""Device input pipeline (pg_prep_batch_u8, patchgan_b200/io.py) against the oracle restatement of io.py:38-58
(oracle/io_oracle.py, itself pinned to the live reference by tests/golden/io.npz): masks bit-exact, image bit-exact against
the oracle (same operation order) and within 1 ulp of the reference's recorded output."""
import os

import numpy as np
import pytest
import torch

import patchgan_b200 as P
from oracle import io_oracle as io
from oracle import patchgan_oracle as orc
from patchgan_b200.io import COCOStuffDataset, DeviceBatches, prepare_batch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_prepare_batch_matches_reference_golden_and_oracle():
    gold = np.load(os.path.join(GOLD, 'io.npz'))
    labels, size = gold['labels'], tuple(int(v) for v in gold['size'])
    for i in range(3):
        iu, lu = gold[f'img_u8/{i}'], gold[f'lab_u8/{i}']
        for code in (0, 1, 2, 3):
            x, y = prepare_batch(torch.from_numpy(iu)[None].cuda(), torch.from_numpy(lu)[None].cuda(), labels, size,
                                 torch.tensor([code], dtype=torch.uint8).cuda())
            ref_img, ref_mask = io.prepare_sample(iu, lu, labels, size, flip=code)
            assert np.array_equal(y[0].cpu().numpy(), ref_mask), (i, code)
            assert np.array_equal(x[0].cpu().numpy(), ref_img), (i, code)
        x, y = prepare_batch(torch.from_numpy(iu)[None].cuda(), torch.from_numpy(lu)[None].cuda(), labels, size)
        assert np.array_equal(y[0].cpu().numpy().astype(np.uint8), gold[f'mask/{i}'])
        assert np.abs(x[0].cpu().numpy() - gold[f'img/{i}']).max() <= 1.2e-7


def test_prepare_batch_large_random_batches():
    r = np.random.default_rng(3)
    for (B, H, W, S) in [(5, 240, 320, (256, 256)), (3, 256, 256, (256, 256)), (2, 101, 57, (128, 64))]:
        iu = r.integers(0, 256, (B, 3, H, W), dtype=np.uint8)
        lu = r.choice(np.array([0, 3, 9, 17, 254, 255], dtype=np.uint8), size=(B, H // 4 + 1, W // 4 + 1)).repeat(4, 1).repeat(4, 2)[:, :H, :W]
        lu = np.ascontiguousarray(lu)
        labels = [4, 10, 0, 18, 255]
        flips = r.integers(0, 4, B).astype(np.uint8)
        x, y = prepare_batch(torch.from_numpy(iu).cuda(), torch.from_numpy(lu).cuda(), labels, S, torch.from_numpy(flips).cuda())
        for b in range(B):
            ri, rm = io.prepare_sample(iu[b], lu[b], labels, S, flip=int(flips[b]))
            assert np.array_equal(y[b].cpu().numpy(), rm)
            assert np.array_equal(x[b].cpu().numpy(), ri)


def test_device_batches_over_files(tmp_path):
    """COCOStuffDataset (raw samples) + DeviceBatches on jpg / png files written here, against the oracle applied to the
    same decoded bytes."""
    from torchvision.io import ImageReadMode, read_image, write_jpeg, write_png
    r = np.random.default_rng(5)
    os.makedirs(tmp_path / 'img'); os.makedirs(tmp_path / 'msk')
    for i, (h, w) in enumerate([(90, 120), (90, 120), (64, 80), (128, 128), (70, 70)]):
        write_jpeg(torch.from_numpy(r.integers(0, 256, (3, h, w), dtype=np.uint8)), str(tmp_path / 'img' / f'{i:04d}.jpg'))
        write_png(torch.from_numpy(r.integers(0, 6, (1, h, w), dtype=np.uint8)), str(tmp_path / 'msk' / f'{i:04d}.png'))
    ds = COCOStuffDataset(str(tmp_path / 'img'), str(tmp_path / 'msk'), labels=[3, 1, 5], size=64, augmentation='randomcrop')
    assert len(ds) == 5
    batches = list(DeviceBatches(ds, batch_size=3, shuffle=False, num_workers=0))
    assert [tuple(b[0].shape) for b in batches] == [(3, 3, 64, 64), (2, 3, 64, 64)]
    k = 0
    for x, y in batches:
        assert x.is_cuda and y.shape[1] == 3
        for b in range(x.shape[0]):
            iu = read_image(ds.images[k], ImageReadMode.RGB).numpy()
            lu = read_image(ds.masks[k], ImageReadMode.GRAY).numpy()[0]
            ri, rm = io.prepare_sample(iu, lu, [3, 1, 5], (64, 64))
            assert np.array_equal(y[b].cpu().numpy(), rm) and np.array_equal(x[b].cpu().numpy(), ri)
            k += 1


def test_trainer_takes_raw_uint8_batches(tmp_path):
    """Trainer.batch(x uint8, label map uint8) == Trainer.batch on the float tensors the reference's dataset would hand over."""
    gk = dict(input_nc=3, output_nc=2, nf=8, activation='tanh', final_act='softmax')
    dk = dict(input_nc=5, ndf=8, n_layers=2, norm=False)
    r = np.random.default_rng(9)
    xu = r.integers(0, 256, (2, 3, 256, 256), dtype=np.uint8)
    lu = r.integers(0, 4, (2, 256, 256), dtype=np.uint8)
    labels = [2, 3]
    xf = np.stack([io.prepare_sample(xu[b], lu[b], labels, (256, 256))[0] for b in range(2)])
    yf = np.stack([io.prepare_sample(xu[b], lu[b], labels, (256, 256))[1] for b in range(2)])
    out = []
    for raw in (False, True):
        og, od = orc.UNet(**gk, seed=11), orc.Discriminator(**dk, seed=12)
        G, D = P.UNet(**gk), P.Discriminator(**dk)
        G.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in og.params.items()})
        D.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in od.params.items()})
        tr = P.Trainer(G.cuda().train(), D.cuda().train(), str(tmp_path / f'c{int(raw)}'))
        tr.loss_type = 'MAE'
        tr.make_optimizers(1e-3, 1e-3)
        if raw:
            tr.labels = labels
            out.append([tr.batch(torch.from_numpy(xu).pin_memory(), torch.from_numpy(lu).pin_memory(), train=True) for _ in range(4)])
        else:
            out.append([tr.batch(torch.from_numpy(xf), torch.from_numpy(yf), train=True) for _ in range(4)])
    for a, b in zip(*out):
        for k in a:
            assert abs(a[k] - b[k]) <= 5e-3 * abs(a[k]), (k, a[k], b[k])
