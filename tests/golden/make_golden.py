"""Generate golden vectors by running the UNMODIFIED reference (imported from
$PATCHGAN_REF, default /root/reference) on CPU.  Run in the build container only:

    python tests/golden/make_golden.py

Writes tests/golden/step_<case>.npz.  The reference cannot travel to the GPU box,
so these fixtures (plus this script) are what pins ``oracle/patchgan_oracle.py``.

Initial weights and inputs are produced by numpy RNG (oracle.UNet / Discriminator
default init, ``synthetic_batch``) and loaded into the reference modules with
``load_state_dict``, so the fixture only needs to hold OUTPUTS: for every tensor
its mean, std and 256 evenly spaced samples.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get('PATCHGAN_REF', '/root/reference')
sys.path.insert(0, REF)
sys.modules.setdefault('torchinfo', types.SimpleNamespace(summary=lambda *a, **k: None))

import torch  # noqa: E402
from torch import optim  # noqa: E402

import patchgan  # noqa: E402  (the reference)
import patchgan.trainer as ref_trainer  # noqa: E402
from oracle import patchgan_oracle as orc  # noqa: E402

ref_trainer.device = 'cpu'

from tests.golden.cases import BIG_CASES, BN_CASES, CASES, NS, rect_batch, summarize  # noqa: E402,F401


def run_case(name, gk, dk, loss_type, B, steps, S=256):
    og = orc.UNet(**gk, seed=11)
    od = orc.Discriminator(**dk, seed=12)
    ref_gk = dict(gk)
    if ref_gk.pop('norm', 'instance') == 'batch':
        ref_gk['norm_layer'] = torch.nn.BatchNorm2d
    G = patchgan.UNet(**ref_gk)
    ref_dk = dict(dk)
    if ref_dk.get('norm_layer') == 'batch':
        ref_dk['norm_layer'] = torch.nn.BatchNorm2d
    D = patchgan.Discriminator(**ref_dk)
    # (strict=False: BatchNorm's running buffers keep their defaults, which are the oracle's too)
    missing = G.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in og.params.items()}, strict=False)
    assert not missing.unexpected_keys and all('running_' in k or 'num_batches' in k for k in missing.missing_keys), missing
    dmiss = D.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in od.params.items()}, strict=False)
    assert not dmiss.unexpected_keys and all('running_' in k or 'num_batches' in k for k in dmiss.missing_keys), dmiss
    import tempfile
    tr = patchgan.Trainer(G, D, tempfile.mkdtemp(), device='cpu')
    tr.loss_type = loss_type
    tr.gen_optimizer = optim.Adam(G.parameters(), lr=1e-3, betas=(0.9, 0.999))    # trainer.py:169-170
    tr.disc_optimizer = optim.Adam(D.parameters(), lr=1e-3, betas=(0.9, 0.999))  # trainer.py:171-172
    G.train()
    D.train()
    out = {}
    acts = {}
    hooks = []

    def grab(key):
        def hook(m, a, o):
            if key not in acts:
                acts[key] = o.detach().numpy().copy()
        return hook
    for i, blk in enumerate(G.encoder):
        hooks.append(blk.register_forward_hook(grab(f'enc{i}')))
    for i, blk in enumerate(G.decoder):
        hooks.append(blk.register_forward_hook(grab(f'dec{i}')))
    ends = [j - 1 for j in od.seq_idx[1:]] + [len(D.model) - 1]
    for li, j in enumerate(ends):
        hooks.append(D.model[j].register_forward_hook(grab(f'd{li}')))
    for step in range(steps):
        x, y = orc.synthetic_batch(B, gk['output_nc'], S, seed=1234 + step)
        losses = tr.batch(torch.from_numpy(x), torch.from_numpy(y), train=True)
        for k, v in losses.items():
            out[f's{step}/loss/{k}'] = np.float64(v)
        if step == 0:
            for k, v in acts.items():
                out[f's0/act/{k}'] = summarize(v)
            for h in hooks:
                h.remove()
        for k, p in G.named_parameters():
            out[f's{step}/ggrad/{k}'] = summarize(p.grad.numpy())
            out[f's{step}/gw/{k}'] = summarize(p.detach().numpy())
        for k, p in D.named_parameters():
            out[f's{step}/dgrad/{k}'] = summarize(p.grad.numpy())
            out[f's{step}/dw/{k}'] = summarize(p.detach().numpy())
        for k, b in G.named_buffers():
            if 'running_' in k:
                out[f's{step}/gbuf/{k}'] = summarize(b.numpy())
        for k, b in D.named_buffers():
            if 'running_' in k:
                out[f's{step}/dbuf/{k}'] = summarize(b.numpy())
    # eval-mode forward / train=False batch (trainer.py:239-259)
    G.eval()
    D.eval()
    x, y = orc.synthetic_batch(B, gk['output_nc'], S, seed=99)
    losses = tr.batch(torch.from_numpy(x), torch.from_numpy(y), train=False)
    for k, v in losses.items():
        out[f'eval/loss/{k}'] = np.float64(v)
    with torch.no_grad():
        out['eval/gen_img'] = summarize(G(torch.from_numpy(x)).numpy())
    np.savez_compressed(os.path.join(HERE, f'step_{name}.npz'), **out)
    print(name, {k: float(v) for k, v in out.items() if '/loss/' in k and k.startswith('s0')})


def rect_case():
    """A rectangular input (128 x 256: the bottleneck is 1 x 2) through a 5-layer discriminator: one training step.
    Pins the oracle's geometry handling where H != W (every other case is square)."""
    gk = dict(input_nc=3, output_nc=1, nf=8, activation='leakyrelu', final_act='sigmoid')
    dk = dict(input_nc=4, ndf=8, n_layers=5, norm=False)
    og, od = orc.UNet(**gk, seed=21), orc.Discriminator(**dk, seed=22)
    G, D = patchgan.UNet(**gk), patchgan.Discriminator(**dk)
    G.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in og.params.items()})
    D.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in od.params.items()})
    import tempfile
    tr = patchgan.Trainer(G, D, tempfile.mkdtemp(), device='cpu')
    tr.loss_type = 'tversky'
    tr.gen_optimizer = optim.Adam(G.parameters(), lr=1e-3, betas=(0.9, 0.999))
    tr.disc_optimizer = optim.Adam(D.parameters(), lr=1e-3, betas=(0.9, 0.999))
    G.train()
    D.train()
    x, y = rect_batch()
    out = {}
    with torch.no_grad():
        out['gen_img'] = summarize(G(torch.from_numpy(x)).numpy())
        out['disc_shape'] = np.array(D(torch.cat([torch.from_numpy(x), torch.from_numpy(y)], 1)).shape)
    losses = tr.batch(torch.from_numpy(x), torch.from_numpy(y), train=True)
    for k, v in losses.items():
        out[f'loss/{k}'] = np.float64(v)
    for k, p in G.named_parameters():
        out[f'ggrad/{k}'] = summarize(p.grad.numpy())
    for k, p in D.named_parameters():
        out[f'dgrad/{k}'] = summarize(p.grad.numpy())
    np.savez_compressed(os.path.join(HERE, 'step_rect.npz'), **out)
    print('rect', {k: float(v) for k, v in out.items() if k.startswith('loss/')}, out['disc_shape'])


def losses_case():
    """Direct calls of losses.py functions (incl. tversky / batch_mean=False)."""
    from patchgan import losses as L
    rng = np.random.default_rng(5)
    p = rng.random((3, 2, 16, 16), dtype=np.float32)
    t = (rng.random((3, 2, 16, 16)) > 0.6).astype(np.float32)
    tp, tt = torch.from_numpy(p), torch.from_numpy(t)
    out = dict(
        tversky=L.tversky(tt, tp, 0.7).item(),
        tversky_nb=L.tversky(tt, tp, 0.7, batch_mean=False).numpy(),
        fc=L.fc_tversky(tt, tp, 0.75, 0.75).item(),
        fc_nb=L.fc_tversky(tt, tp, 0.75, 0.75, batch_mean=False).numpy(),
        mae=L.MAE_loss(tt, tp).item(),
        bce=L.bce_loss(tp, tt).item(),
    )
    # gradients wrt the prediction of every function / mode (autograd of the reference); per-sample results are
    # reduced with the weights (1, 2, 3) so that each sample's gradient is told apart
    wv = torch.tensor([1., 2., 3.])
    fns = dict(tversky=lambda q: L.tversky(tt, q, 0.7), tversky_nb=lambda q: (L.tversky(tt, q, 0.7, batch_mean=False) * wv).sum(),
               fc=lambda q: L.fc_tversky(tt, q, 0.75, 0.75),
               fc_nb=lambda q: (L.fc_tversky(tt, q, 0.75, 0.75, batch_mean=False) * wv).sum(),
               mae=lambda q: L.MAE_loss(tt, q), bce=lambda q: L.bce_loss(q, tt))
    for k, f in fns.items():
        q = tp.clone().requires_grad_(True)
        f(q).backward()
        out['grad_' + k] = q.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, 'losses.npz'), **out)


def io_case():
    """COCOStuffDataset.__getitem__ (io.py:38-58) on synthetic jpg / png files.  The decoded uint8 arrays are stored with the
    outputs, so the tests need neither an image decoder nor the files; flip codes are detected from the output."""
    import tempfile
    from torchvision.io import ImageReadMode, read_image, write_jpeg, write_png
    from patchgan.io import COCOStuffDataset
    rng = np.random.default_rng(17)
    d = tempfile.mkdtemp()
    os.makedirs(d + '/img'); os.makedirs(d + '/msk')
    sizes = [(75, 100), (64, 64), (130, 97)]
    for i, (h, w) in enumerate(sizes):
        # smooth-ish image, blocky label map with values that include 254 / 255 (uint8 wrap of the +1 shift)
        img = (rng.random((3, h // 4 + 1, w // 4 + 1)) * 255).astype(np.uint8).repeat(4, 1).repeat(4, 2)[:, :h, :w]
        img = (img.astype(np.int32) + rng.integers(-8, 9, img.shape)).clip(0, 255).astype(np.uint8)
        lab = rng.choice(np.array([0, 1, 2, 6, 254, 255], dtype=np.uint8), size=(h // 8 + 1, w // 8 + 1)).repeat(8, 0).repeat(8, 1)[:h, :w]
        write_jpeg(torch.from_numpy(np.ascontiguousarray(img)), f'{d}/img/{i:03d}.jpg', quality=95)
        write_png(torch.from_numpy(np.ascontiguousarray(lab[None])), f'{d}/msk/{i:03d}.png')
    labels = [7, 1, 3, 0]                      # unsorted on purpose (io.py:17 sorts); 0 = wrapped 255
    S = 64
    out = dict(labels=np.array(labels), size=np.array([S, S]))
    ds = COCOStuffDataset(d + '/img', d + '/msk', labels=labels, size=S, augmentation='randomcrop')
    for i in range(len(sizes)):
        out[f'img_u8/{i}'] = read_image(ds.images[i], ImageReadMode.RGB).numpy()
        out[f'lab_u8/{i}'] = read_image(ds.masks[i], ImageReadMode.GRAY).numpy()[0]
        img, mask = ds[i]
        out[f'img/{i}'] = img.numpy()
        out[f'mask/{i}'] = mask.numpy().astype(np.uint8)
    # flips: run the flipping dataset under several seeds, keep one sample per flip code that occurred
    dsf = COCOStuffDataset(d + '/img', d + '/msk', labels=labels, size=S, augmentation='randomcrop+flip')
    seen = {}
    for seed in range(200):
        torch.manual_seed(seed)
        img, mask = dsf[0]
        base = out['img/0']
        for code in (1, 2, 3):
            ref = base[:, :, ::-1] if code & 1 else base
            ref = ref[:, ::-1, :] if code & 2 else ref
            if code not in seen and np.array_equal(img.numpy(), ref):
                seen[code] = (img.numpy(), mask.numpy().astype(np.uint8))
        if len(seen) == 3:
            break
    for code, (img, mask) in seen.items():
        out[f'flip{code}/img'] = img
        out[f'flip{code}/mask'] = mask
    out['flip_codes'] = np.array(sorted(seen))
    np.savez_compressed(os.path.join(HERE, 'io.npz'), **out)
    print('io', {k: v.shape for k, v in out.items() if k.startswith('img/')}, 'flip codes', sorted(seen))


def infer_case():
    """n_crop / build_mask (infer.py:14-68) on a small non-trivial image."""
    from patchgan import infer as I
    rng = np.random.default_rng(7)
    img = rng.random((3, 300, 300), dtype=np.float32)
    crops = I.n_crop(torch.from_numpy(img), 128, 0.9).numpy()
    masks = rng.random((crops.shape[0], 4, 128, 128), dtype=np.float32)
    m_arg = I.build_mask(masks, 128, (300, 300), 0.0, 0.9)
    m_thr = I.build_mask(masks[:, :1], 128, (300, 300), 0.5, 0.9)
    np.savez_compressed(os.path.join(HERE, 'infer.npz'), crops_sum=summarize(crops),
                        m_arg=m_arg.astype(np.int16), m_thr=m_thr.astype(np.float32))


if __name__ == '__main__':
    torch.manual_seed(0)
    torch.set_num_threads(8)
    only = sys.argv[1:]               # e.g. `make_golden.py rect` regenerates one fixture
    for name, (gk, dk, lt, B, steps) in CASES.items():
        if not only or name in only:
            run_case(name, gk, dk, lt, B, steps)
    for name, (gk, dk, lt, B, S, steps) in BIG_CASES.items():
        if not only or name in only:
            run_case(name, gk, dk, lt, B, steps, S)
    for name, (gk, dk, lt, B, steps) in BN_CASES.items():
        if not only or name in only:
            run_case(name, gk, dk, lt, B, steps)
    if not only or 'rect' in only:
        rect_case()
    if not only or 'losses' in only:
        losses_case()
    if not only or 'infer' in only:
        infer_case()
    if not only or 'io' in only:
        io_case()
