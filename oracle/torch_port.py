"""CPU baseline port of the reference step on the reference's OWN arithmetic library -- TEST / BENCH
INFRASTRUCTURE, NOT PRODUCT CODE (same import rules as patchgan_oracle.py).

The reference (/root/reference/patchgan) delegates all arithmetic to PyTorch's CPU operators.  The numpy oracle in
patchgan_oracle.py is the independent checker, but it is ~15x slower than those operators, which would flatter any
GPU/CPU ratio.  This module restates the same step -- UNet.forward (unet.py:112-134), Discriminator.forward
(disc.py:49-51), Trainer.batch (trainer.py:50-115) with optim.Adam (trainer.py:169-172) -- as a functional program
over ``torch.nn.functional`` and autograd, so that ``bench.py --impl reference`` and the ``cpu_baseline`` leg time
what the reference would cost on the same host cores.  It is pinned against the live reference's golden vectors in
tests/test_oracle_golden.py::test_torch_port_matches_reference.
"""
import torch
import torch.nn.functional as F


def _act(name, x):
    if name == 'tanh':
        return torch.tanh(x)
    if name == 'relu':
        return F.relu(x)
    if name == 'leakyrelu':
        return F.leaky_relu(x, 0.2)
    if name == 'sigmoid':
        return torch.sigmoid(x)
    if name == 'softmax':
        return torch.softmax(x, dim=1)
    raise ValueError(name)


def unet_forward(p, x, activation, final_act, training=False, use_dropout=False):
    """p: dict of reference-named tensors (encoder.{i}.model.DownConv{i}.weight, decoder.{i}.model.UpConv{i}.weight)."""
    encs = []
    h = x
    for i in range(7):
        h = F.conv2d(h, p[f'encoder.{i}.model.DownConv{i}.weight'], None, stride=2, padding=1)
        h = _act(activation, F.instance_norm(h, eps=1e-5))
        if use_dropout:
            h = F.dropout(h, 0.2, training)
        encs.append(h)
    skips = encs[::-1]
    for i in range(7):
        inp = h if i == 0 else torch.cat([h, skips[i]], dim=1)
        h = F.conv_transpose2d(inp, p[f'decoder.{i}.model.UpConv{i}.weight'], None, stride=2, padding=1)
        if 1 <= i <= 5:
            h = _act(activation, F.instance_norm(h, eps=1e-5))
            if use_dropout:
                h = F.dropout(h, 0.2, training)
        else:
            h = _act(final_act if i == 6 else activation, h)
    return h


def disc_layout(n_layers, norm):
    """(sequential index, stride, has_bias, act) per conv of disc.py:19-46."""
    out, idx = [(0, 2, True, 'leakyrelu')], 2
    for _ in range(1, n_layers):
        out.append((idx, 2, False, 'tanh'))
        idx += 3 if norm else 2
    out.append((idx, 1, False, 'tanh'))
    idx += 3 if norm else 2
    out.append((idx, 1, True, 'sigmoid'))
    return out


def disc_forward(p, x, n_layers, norm=False):
    h = x
    lay = disc_layout(n_layers, norm)
    for li, (k, s, b, a) in enumerate(lay):
        h = F.conv2d(h, p[f'model.{k}.weight'], p[f'model.{k}.bias'] if b else None, stride=s, padding=1)
        h = _act(a, h)
        if norm and 0 < li < len(lay) - 1:
            h = F.instance_norm(h, eps=1e-5)
    return h


def seg_loss(loss_type, target, gen_img, beta=0.75, gamma=0.75):
    if loss_type == 'tversky':        # losses.py:18-31
        tp = torch.sum(target * gen_img, dim=(1, 2, 3))
        fn = torch.sum((1. - gen_img) * target, dim=(1, 2, 3))
        fp = torch.sum(gen_img * (1. - target), dim=(1, 2, 3))
        tv = (tp + 1) / (tp + beta * fn + (1. - beta) * fp + 1)
        return torch.pow(torch.mean(1 - tv), gamma)
    if loss_type == 'weighted_bce':   # trainer.py:75-80
        if gen_img.shape[1] > 1:
            w = 1 - torch.sum(target, dim=(2, 3), keepdim=True) / torch.sum(target)
        else:
            w = torch.ones_like(target)
        return F.binary_cross_entropy(gen_img, target, weight=w)
    if loss_type == 'MAE':            # losses.py:34-35
        return torch.mean(torch.abs(gen_img - target))
    raise ValueError(loss_type)


class Step:
    """Trainer.batch(train=True) (trainer.py:50-115) over functional nets; parameters are leaf tensors."""

    def __init__(self, gparams, dparams, gcfg, dcfg, loss_type='tversky', seg_alpha=200, lr=1e-3):
        self.g = {k: torch.as_tensor(v).clone().requires_grad_(True) for k, v in gparams.items()}
        self.d = {k: torch.as_tensor(v).clone().requires_grad_(True) for k, v in dparams.items()}
        self.gcfg, self.dcfg, self.loss_type, self.seg_alpha = gcfg, dcfg, loss_type, seg_alpha
        self.gopt = torch.optim.Adam(list(self.g.values()), lr=lr, betas=(0.9, 0.999))
        self.dopt = torch.optim.Adam(list(self.d.values()), lr=lr, betas=(0.9, 0.999))

    def G(self, x, training=True):
        c = self.gcfg
        return unet_forward(self.g, x, c['activation'], c['final_act'], training, c.get('use_dropout', False))

    def D(self, x):
        return disc_forward(self.d, x, self.dcfg['n_layers'], self.dcfg.get('norm', False))

    def batch(self, x, y, train=True):
        gen_img = self.G(x, train)
        disc_fake = self.D(torch.cat((x, gen_img), 1))
        ones, zeros = torch.ones_like(disc_fake), torch.zeros_like(disc_fake)
        gen_loss_disc = F.binary_cross_entropy(disc_fake, ones)
        gen_loss = seg_loss(self.loss_type, y, gen_img) * self.seg_alpha + gen_loss_disc
        if train:
            self.gopt.zero_grad()
            gen_loss.backward()            # also fills (and wastes) the discriminator's gradients, like the reference
            self.gopt.step()
            self.dopt.zero_grad()
        disc_real = self.D(torch.cat((x, y), 1))
        disc_fake = self.D(torch.cat((x, gen_img.detach()), 1))
        loss_real = F.binary_cross_entropy(disc_real, ones)
        loss_fake = F.binary_cross_entropy(disc_fake, zeros)
        disc_loss = (loss_fake + loss_real) / 2.
        if train:
            disc_loss.backward()
            self.dopt.step()
        vals = [gen_loss.item(), gen_loss.item(), gen_loss_disc.item(), loss_real.item(), loss_fake.item(),
                disc_loss.item()]
        return dict(zip(['gen', 'gen_loss', 'gdisc', 'discr', 'discf', 'disc'], vals))
