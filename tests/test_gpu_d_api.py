"""GPU tests of the reference-facing Python API beyond a single Trainer.batch: the public loss functions (values and
gradients against goldens recorded from the live reference's losses.py + autograd), Trainer.train() with a learning-rate
schedule under CUDA-graph replay, re-creating the optimizers after a graph was captured, optimizer-state checkpoints,
the differentiable encoder bottleneck of UNet.forward(return_hidden=True), and the `patchgan` package alias."""
import os

import numpy as np
import pytest
import torch

import patchgan_b200 as P
from oracle import patchgan_oracle as orc
from patchgan_b200 import losses as PL
from tests.golden.cases import CASES
from tests.gpu_util import relerr

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def loss_inputs():
    rng = np.random.default_rng(5)
    p = rng.random((3, 2, 16, 16), dtype=np.float32)
    t = (rng.random((3, 2, 16, 16)) > 0.6).astype(np.float32)
    return torch.from_numpy(p).cuda(), torch.from_numpy(t).cuda()


def test_public_loss_functions_match_reference_values_and_gradients():
    """losses.py:5-39 through patchgan_b200.losses: every function and mode, forward value (1e-5 rel) and gradient wrt the
    prediction (1e-2 norm-wise: the gradient kernels write bf16) against tests/golden/losses.npz."""
    gold = np.load(os.path.join(GOLD, 'losses.npz'))
    p, t = loss_inputs()
    wv = torch.tensor([1., 2., 3.], device='cuda')
    fns = dict(tversky=lambda q: PL.tversky(t, q, 0.7), tversky_nb=lambda q: PL.tversky(t, q, 0.7, batch_mean=False),
               fc=lambda q: PL.fc_tversky(t, q, 0.75, 0.75), fc_nb=lambda q: PL.fc_tversky(t, q, 0.75, 0.75, batch_mean=False),
               mae=lambda q: PL.MAE_loss(t, q), bce=lambda q: PL.bce_loss(q, t))
    for k, f in fns.items():
        q = p.clone().requires_grad_(True)
        val = f(q)
        ref = gold[k]
        assert np.allclose(val.detach().cpu().numpy(), ref, rtol=2e-5, atol=0), (k, val, ref)
        (val if val.dim() == 0 else (val * wv).sum()).backward()
        assert q.grad is not None and q.grad.shape == p.shape, k
        err = relerr(q.grad.cpu().numpy(), gold['grad_' + k])
        assert err < 1e-2, (k, err)


def test_bce_loss_needs_no_host_sync_and_takes_general_targets():
    p, t = loss_inputs()
    soft = torch.rand_like(p)
    q = p.clone().requires_grad_(True)
    loss = PL.bce_loss(q, soft)
    ref = torch.nn.functional.binary_cross_entropy(p.cpu(), soft.cpu())
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    (3.0 * loss).backward()
    pc = p.cpu().clone().requires_grad_(True)
    (3.0 * torch.nn.functional.binary_cross_entropy(pc, soft.cpu())).backward()
    assert relerr(q.grad.cpu().numpy(), pc.grad.numpy()) < 1e-5


def build(tmp_path, name='mae', seeds=(11, 12)):
    gk, dk, loss_type, B, steps = CASES[name]
    og, od = orc.UNet(**gk, seed=seeds[0]), orc.Discriminator(**dk, seed=seeds[1])
    G, D = P.UNet(**gk), P.Discriminator(**dk)
    G.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in og.params.items()})
    D.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in od.params.items()})
    tr = P.Trainer(G.cuda().train(), D.cuda().train(), str(tmp_path / 'ckpt'))
    tr.loss_type = loss_type
    return tr, gk, B


def batches(gk, B, n, seed0=300):
    out = []
    for i in range(n):
        x, y = orc.synthetic_batch(B, gk['output_nc'], 256, seed=seed0 + i)
        out.append((torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()))
    return out


@pytest.mark.timeout(300)
def test_train_epochs_with_lr_decay_under_graph_replay(tmp_path, capsys):
    """Trainer.train (trainer.py:117-279) for three epochs with ExponentialLR: the learning rate the Adam KERNEL reads
    (device memory, read by the replayed graph) follows the schedule, checkpoints appear, losses are returned per epoch."""
    tr, gk, B = build(tmp_path)
    data = batches(gk, B, 4)
    G_loss, D_loss = tr.train(data, data[:2], 3, dsc_learning_rate=2e-3, gen_learning_rate=1e-3, save_freq=2, lr_decay=0.5,
                              decay_freq=1)
    assert len(G_loss) == 3 and len(D_loss) == 3 and all(np.isfinite(G_loss)) and all(np.isfinite(D_loss))
    assert any(e['graph'] is not None for e in tr._graphs.values())                # the epochs ran on replays
    # two decays have been applied to the device-side lr by the time the third epoch ran (lr 1e-3 -> 2.5e-4)
    assert abs(tr.gen_optimizer.flat()['hyper'][0].item() - 2.5e-4) < 1e-9
    assert abs(tr.disc_optimizer.flat()['hyper'][0].item() - 5e-4) < 1e-9
    assert int(tr.gen_optimizer.flat()['step'].item()) == 12
    assert os.path.exists(tr.savefolder + 'generator_ep_002.pth') and os.path.exists(tr.savefolder + 'optimizer_ep_002.pth')
    # the size of a late Adam update is bounded by the DECAYED lr (|update| <= ~lr per step for Adam)
    w0 = {k: p.detach().clone() for k, p in tr.generator.named_parameters()}
    tr.batch(*data[0], train=True)
    step = max(float((p.detach() - w0[k]).abs().max()) for k, p in tr.generator.named_parameters())
    assert 0 < step <= 2.5e-4 * 3.5, step


@pytest.mark.timeout(300)
def test_new_optimizers_after_graph_capture_keep_training_the_live_weights(tmp_path):
    """A second make_optimizers() (every Trainer.train call makes one) re-homes the parameters into new flat buffers.  The
    graphs captured before must not be replayed: the live weights keep changing and follow the eager trajectory."""
    tr, gk, B = build(tmp_path)
    tr_ref, _, _ = build(tmp_path)
    tr_ref.use_cuda_graph = False
    data = batches(gk, B, 9)
    for t in (tr, tr_ref):
        t.make_optimizers(1e-3, 1e-3)
    a = [tr.batch(x, y, train=True) for x, y in data[:5]]                          # past GRAPH_WARMUP: graph captured
    b = [tr_ref.batch(x, y, train=True) for x, y in data[:5]]
    assert any(e['graph'] is not None for e in tr._graphs.values())
    for t in (tr, tr_ref):
        t.make_optimizers(1e-3, 1e-3)                                              # moments restart, new flat buffers
    assert not tr._graphs
    w0 = {k: p.detach().clone() for k, p in tr.generator.named_parameters()}
    a += [tr.batch(x, y, train=True) for x, y in data[5:]]
    b += [tr_ref.batch(x, y, train=True) for x, y in data[5:]]
    assert all(float((p.detach() - w0[k]).abs().max()) > 0 for k, p in tr.generator.named_parameters())
    for u, v in zip(a, b):
        for k in u:
            assert abs(u[k] - v[k]) <= 5e-3 * abs(v[k]), (k, u[k], v[k])
    # four Adam steps after the restart: a weight whose gradient sign differs between the two runs (atomics order) moves
    # apart by up to 2 lr per step; almost all weights agree closely
    for (k, p), (_, q) in zip(tr.generator.named_parameters(), tr_ref.generator.named_parameters()):
        diff = (p - q).abs()
        assert float(diff.max()) <= 4 * 2.05e-3, k
        assert float(diff.mean()) <= 2e-4, (k, float(diff.mean()))


def test_optimizer_state_is_checkpointed_and_resumed(tmp_path):
    tr, gk, B = build(tmp_path)
    tr.make_optimizers(1e-3, 2e-3)
    data = batches(gk, B, 3)
    for x, y in data:
        tr.batch(x, y, train=True)
    tr.save(7)
    tr2, _, _ = build(tmp_path, seeds=(77, 78))
    tr2.load_last_checkpoint()
    assert tr2.start == 8
    tr2.make_optimizers(1e-3, 2e-3)
    for o1, o2 in ((tr.gen_optimizer, tr2.gen_optimizer), (tr.disc_optimizer, tr2.disc_optimizer)):
        f1, f2 = o1.flat(), o2.flat()
        assert int(f2['step'].item()) == 3
        assert torch.equal(f1['m'], f2['m']) and torch.equal(f1['v'], f2['v']) and torch.equal(f1['p'], f2['p'])
    # and both continue identically
    x, y = data[0]
    l1, l2 = tr.batch(x, y, train=True), tr2.batch(x, y, train=True)
    for k in l1:
        assert abs(l1[k] - l2[k]) <= 1e-4 * abs(l1[k]) + 1e-7, (k, l1[k], l2[k])


def test_return_hidden_is_differentiable_and_from_the_same_pass():
    gk = dict(input_nc=3, output_nc=1, nf=8, activation='leakyrelu', final_act='sigmoid')
    og = orc.UNet(**gk, seed=3)
    G = P.UNet(**gk)
    G.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in og.params.items()})
    G = G.cuda().train()
    x, _ = orc.synthetic_batch(2, 1, 256, seed=5)
    out, hidden = G(torch.from_numpy(x).cuda(), return_hidden=True)
    og.forward(x)
    assert hidden.shape == (2, 64, 2, 2) and hidden.requires_grad
    assert relerr(hidden.detach().cpu().numpy(), og.acts['enc6']) < 1e-2
    # a loss on the bottleneck alone reaches the encoder weights and not the decoder's
    m = np.random.default_rng(4).standard_normal(hidden.shape).astype(np.float32)
    (hidden * torch.from_numpy(m).cuda()).sum().backward()
    ps = dict(G.named_parameters())
    assert float(ps['encoder.3.model.DownConv3.weight'].grad.abs().max()) > 0
    assert float(ps['decoder.2.model.UpConv2.weight'].grad.abs().max()) == 0
    # directional derivative of sum(hidden * m) wrt one encoder weight: central difference on the ORACLE (fp32 numpy,
    # no 16-bit noise) against the CUDA path's analytic gradient
    name = 'encoder.5.model.DownConv5.weight'
    g = ps[name].grad.detach().cpu().numpy()
    d = np.random.default_rng(6).standard_normal(g.shape).astype(np.float32)
    d /= np.linalg.norm(d)
    eps = 1e-2
    w0 = og.params[name].copy()
    vals = []
    for sgn in (+1, -1):
        og.params[name] = w0 + sgn * eps * d
        og.forward(x)
        vals.append(float((og.acts['enc6'].astype(np.float64) * m).sum()))
    og.params[name] = w0
    fd = (vals[0] - vals[1]) / (2 * eps)
    an = float((g.astype(np.float64) * d).sum())
    assert abs(fd - an) <= 0.15 * max(abs(fd), abs(an)) + 1e-3, (fd, an)      # (LeakyReLU kinks: the FD itself moves ~5 % with eps)


def test_patchgan_alias_package_and_transfer_loading():
    import patchgan
    from patchgan.losses import fc_tversky  # noqa: F401
    from patchgan.transfer import InvalidCheckpointError
    assert patchgan.UNet is P.UNet and patchgan.Trainer is P.Trainer and patchgan.trainer.Trainer is P.Trainer
    G = P.UNet(3, 1, 8, activation='leakyrelu', final_act='sigmoid').cuda()
    src = P.UNet(3, 2, 8, activation='leakyrelu', final_act='sigmoid').cuda()      # last layer has another shape
    before = G.decoder[6].model.UpConv6.weight.detach().clone()
    ptr = G.encoder[0].model.DownConv0.weight.data_ptr()
    G.load_transfer_data(src.state_dict())
    assert torch.equal(G.encoder[2].model.DownConv2.weight, src.encoder[2].model.DownConv2.weight)
    assert torch.equal(G.decoder[6].model.UpConv6.weight, before)                  # shape mismatch: left alone
    assert G.encoder[0].model.DownConv0.weight.data_ptr() == ptr                   # storage kept (optimizer views stay valid)
    with pytest.raises(InvalidCheckpointError):
        G.load_transfer_data({'nothing.matches': torch.zeros(3)})
