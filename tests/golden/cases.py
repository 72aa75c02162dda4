"""Configurations and the summary statistic shared by make_golden.py (which needs the reference) and the tests
(which must not)."""
import numpy as np

CASES = {
    # name: (G kwargs, D kwargs, loss_type, B, steps)
    'tversky': (dict(input_nc=3, output_nc=1, nf=8, activation='leakyrelu', final_act='sigmoid'),
                dict(input_nc=4, ndf=8, n_layers=3, norm=False), 'tversky', 2, 2),
    'wbce': (dict(input_nc=3, output_nc=3, nf=8, activation='relu', final_act='sigmoid'),
             dict(input_nc=6, ndf=8, n_layers=4, norm=True), 'weighted_bce', 2, 2),
    'mae': (dict(input_nc=3, output_nc=2, nf=8, activation='tanh', final_act='softmax'),
            dict(input_nc=5, ndf=8, n_layers=2, norm=False), 'MAE', 2, 2),
}
NS = 256

# norm_layer = nn.BatchNorm2d in the generator (unet.py:77; the discriminator keeps its default norm=False).  'norm': 'batch' is
# the oracle's spelling; make_golden.py / the GPU tests translate it to norm_layer=nn.BatchNorm2d.  Fixture: step_bn.npz
BN_CASES = {
    'bn': (dict(input_nc=3, output_nc=1, nf=8, activation='leakyrelu', final_act='sigmoid', norm='batch'),
           dict(input_nc=4, ndf=8, n_layers=3, norm=False), 'tversky', 2, 2),
    # BatchNorm2d in BOTH networks: the discriminator's three calls per step (fake, real, fake again: trainer.py:65,96,98)
    # use separate batch statistics and update its running buffers three times.  Fixture: step_bnd.npz
    'bnd': (dict(input_nc=3, output_nc=1, nf=8, activation='tanh', final_act='sigmoid', norm='batch'),
            dict(input_nc=4, ndf=8, n_layers=3, norm=True, norm_layer='batch'), 'MAE', 2, 2),
}

# The BASELINE.json architectures at (or near) their benchmarked sizes -- goldens from the live reference only (the numpy
# oracle would take minutes at these sizes; it is pinned by the small cases above):
#   name: (G kwargs, D kwargs, loss_type, B, S, steps)
BIG_CASES = {
    # cfg 1 / 3: UNet(3 -> 1, nf=32) + PatchGAN(ndf=64, L=3) at the benchmarked batch 16 (D runs at 2B = 32)
    'cfg3_b16': (dict(input_nc=3, output_nc=1, nf=32, activation='leakyrelu', final_act='sigmoid'),
                 dict(input_nc=4, ndf=64, n_layers=3, norm=False), 'tversky', 16, 256, 2),
    # cfg 4: train_coco.yaml-shaped (3 -> 7, ReLU, ndf=16, L=5, weighted BCE); dropout off (RNG streams cannot match)
    'cfg4_b4': (dict(input_nc=3, output_nc=7, nf=32, activation='relu', final_act='sigmoid'),
                dict(input_nc=10, ndf=16, n_layers=5, norm=False), 'weighted_bce', 4, 256, 2),
    # cfg 5: wide generator nf=64 + 4-layer PatchGAN at 1024 x 1024, one image
    'cfg5_b1': (dict(input_nc=3, output_nc=1, nf=64, activation='leakyrelu', final_act='sigmoid'),
                dict(input_nc=4, ndf=64, n_layers=4, norm=False), 'tversky', 1, 1024, 1),
}


def summarize(t):
    a = np.asarray(t, dtype=np.float64).ravel()
    idx = np.linspace(0, a.size - 1, min(NS, a.size)).astype(np.int64)
    return np.concatenate([[a.mean(), a.std()], a[idx]])




def rect_batch(B=2, H=128, W=256, seed=4321):
    """Rectangular batch of the 'rect' fixture: x in [0,1), binary one-channel target."""
    rng = np.random.default_rng(seed)
    x = rng.random((B, 3, H, W), dtype=np.float32)
    y = (rng.random((B, 1, H, W), dtype=np.float32) > 0.5).astype(np.float32)
    return x, y
