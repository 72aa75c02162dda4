"""CPU-only checks of the host side: the C-ABI library loads and exports every declared symbol, the modules keep
the reference's constructor / state_dict contract, and the product fails loudly without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from patchgan_b200 import _lib as L
    from patchgan_b200.build import build
    build()
    header = open(os.path.join(ROOT, 'include', 'patchgan_b200.h')).read()
    declared = set(re.findall(r'\b(pg_[a-z0-9_]+)\s*\(', header))
    assert len(declared) >= 24
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/patchgan_b200.h but not exported'
    assert declared == set(L.exported_symbols()), declared ^ set(L.exported_symbols())
    assert L.lib().pg_version() == 100


def test_no_cpu_fallback():
    import patchgan_b200 as P
    G = P.UNet(3, 1, 8, activation='leakyrelu', final_act='sigmoid')
    with pytest.raises(RuntimeError, match='CUDA'):
        G(torch.zeros(1, 3, 256, 256))
    D = P.Discriminator(4, 8)
    with pytest.raises(RuntimeError, match='CUDA'):
        D(torch.zeros(1, 4, 256, 256))
    from patchgan_b200.losses import fc_tversky
    with pytest.raises(RuntimeError, match='CUDA'):
        fc_tversky(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4, 4), 0.75)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'patchgan_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), fn


def test_state_dict_contract_matches_reference_layout():
    import patchgan_b200 as P
    from oracle import patchgan_oracle as orc
    for gk in (dict(input_nc=3, output_nc=1, nf=32), dict(input_nc=4, output_nc=7, nf=8)):
        G = P.UNet(**gk)
        og = orc.UNet(**gk)
        sd = G.state_dict()
        assert list(sd) == og.enc_names + og.dec_names
        for k in sd:
            assert tuple(sd[k].shape) == og.params[k].shape and sd[k].dtype == torch.float32
    for dk in (dict(input_nc=4, ndf=64, n_layers=3), dict(input_nc=10, ndf=16, n_layers=5, norm=True)):
        D = P.Discriminator(**dk)
        od = orc.Discriminator(**dk)
        assert list(D.state_dict()) == list(od.params)
        for k, v in D.state_dict().items():
            assert tuple(v.shape) == od.params[k].shape
    # default init is U(+-1/sqrt(fan_in)) like nn.Conv2d (trainer.py:327-343 weights_init is a no-op)
    w = P.UNet(3, 1, 32).state_dict()['encoder.1.model.DownConv1.weight']
    assert float(w.abs().max()) <= 1 / (32 * 16) ** 0.5 + 1e-7 and float(w.abs().max()) > 0.9 / (32 * 16) ** 0.5


def test_transfer_learning_partial_load():
    import patchgan_b200 as P
    from patchgan_b200.transfer import InvalidCheckpointError
    a, b = P.UNet(3, 1, 8), P.UNet(3, 2, 8)
    b.load_transfer_data(a.state_dict())            # all but the last layer match
    assert torch.equal(a.state_dict()['encoder.0.model.DownConv0.weight'],
                       b.state_dict()['encoder.0.model.DownConv0.weight'])
    with pytest.raises(InvalidCheckpointError):
        P.UNet(3, 1, 16).load_transfer_data({'encoder.0.model.DownConv0.weight': torch.zeros(1)})


def test_yaml_schemas_both_accepted():
    """train.py reads the nested schema, infer.py / examples/train_coco.yaml the flat one."""
    from patchgan_b200 import config as cfg
    nested = {'model_params': {'generator': {'filters': 32, 'activation': 'relu'},
                               'discriminator': {'filters': 16, 'n_layers': 5}},
              'dataset': {'type': 'COCOStuff', 'labels': [1, 2, 3], 'train_data': {'images': 'a', 'masks': 'b'},
                          'validation_data': {'images': 'c', 'masks': 'd'}}}
    flat = {'model_params': {'gen_filts': 32, 'disc_filts': 16, 'activation': 'relu', 'use_dropout': True,
                             'final_activation': 'sigmoid', 'n_disc_layers': 5},
            'dataset': {'type': 'COCOStuff'}, 'train_data': {'images': 'a', 'masks': 'b'},
            'validation_data': {'images': 'c', 'masks': 'd'}}
    a, b = cfg.model_params(nested), cfg.model_params(flat)
    assert a == b == dict(gen_filts=32, activation='relu', use_dropout=True, final_activation='sigmoid',
                          disc_filts=16, disc_norm=False, n_disc_layers=5)
    assert cfg.data_paths(nested)[0] == cfg.data_paths(flat)[0] == {'images': 'a', 'masks': 'b'}
    cls, cin, cout, kw = cfg.dataset_class(nested['dataset'])
    assert (cls.__name__, cin, cout, kw) == ('COCOStuffDataset', 3, 3, {'labels': [1, 2, 3]})
    with pytest.raises(AttributeError):
        cfg.data_paths({'dataset': {'type': 'COCOStuff'}})


def test_console_entry_points_importable():
    from patchgan_b200.infer import build_mask, n_crop, patchgan_infer  # noqa: F401
    from patchgan_b200.train import patchgan_train  # noqa: F401
    with pytest.raises(RuntimeError, match='CUDA'):
        n_crop(torch.zeros(3, 300, 300), 128, 0.9)


def test_pending_losses_builds_the_reference_loss_dict():
    """Trainer.submit's handle: waits on its event once, then returns the six-entry dict of trainer.py:109-113
    (gen = gen_loss = seg*alpha + gdisc in fp32, disc = (discf + discr) / 2)."""
    import numpy as np
    from patchgan_b200.trainer import LOSS_KEYS, PendingLosses

    class Ev:
        def __init__(self):
            self.waits = 0

        def synchronize(self):
            self.waits += 1

        def query(self):
            return False

    host = torch.tensor([55.5, 7.25, 0.5, 0.25, 0, 0, 0, 0], dtype=torch.float32)
    ev = Ev()
    h = PendingLosses(host, ev)
    assert not h.done()
    d = h.result()
    assert list(d) == LOSS_KEYS == ['gen', 'gen_loss', 'gdisc', 'discr', 'discf', 'disc']
    assert d['gen'] == d['gen_loss'] == float(np.float32(55.5) + np.float32(7.25))
    assert (d['gdisc'], d['discr'], d['discf']) == (7.25, 0.5, 0.25)
    assert d['disc'] == 0.375
    host[0] = 1.0                       # the pinned slot is recycled by a later submit: the handle keeps its values
    assert h.result() is d and h.done() and ev.waits == 1


def test_integration_stub_structs_match_the_binding():
    """The ctypes structures shown in INTEGRATION.md (what a maintainer of the reference would paste) have the layout of the
    ones this package binds with (patchgan_b200/_lib.py), which the C side checks with static_asserts."""
    import ctypes as C
    import re
    from patchgan_b200 import _lib as L
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    blocks = re.findall(r'^class (Pg\w+)\(C\.Structure\):.*?(?=^\S|\Z)', text, flags=re.S | re.M)
    assert set(blocks) >= {'PgConvDesc', 'PgFusedNorm'}, blocks
    ns = {'C': C}
    for name in ('PgConvDesc', 'PgFusedNorm'):
        src = re.search(rf'^class {name}\(C\.Structure\):.*?(?=^\S)', text, flags=re.S | re.M).group(0)
        exec(src, ns)
    for stub, ours in ((ns['PgConvDesc'], L.ConvDesc), (ns['PgFusedNorm'], L.FusedNorm)):
        assert C.sizeof(stub) == C.sizeof(ours), (stub, C.sizeof(stub), C.sizeof(ours))
        assert [(n, getattr(stub, n).offset) for n, _ in stub._fields_] == \
               [(n, getattr(ours, n).offset) for n, _ in ours._fields_], stub
