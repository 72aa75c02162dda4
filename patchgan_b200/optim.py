"""Fused multi-tensor Adam: ONE kernel launch updates every parameter of a network.

Semantics are ``torch.optim.Adam(params, lr, betas=(0.9, 0.999))`` with eps=1e-8, weight_decay=0, no amsgrad --
exactly what the reference constructs at /root/reference/patchgan/trainer.py:169-172 and steps at :90 / :107.
The class subclasses ``torch.optim.Optimizer`` so torch LR schedulers (ExponentialLR / ReduceLROnPlateau,
trainer.py:175-188) drive ``param_groups[0]['lr']`` unchanged.

To make the update a single launch, the parameters are re-homed into one flat fp32 buffer (each
``nn.Parameter.data`` becomes a view of it, each ``.grad`` a view of the flat gradient buffer).  The learning rate
and the step count live in device memory so a captured CUDA graph replays correctly when the scheduler changes lr.
"""
import ctypes

import torch

from . import _lib as L
from .engine import _stream


def _align4(n):
    return (n + 3) // 4 * 4


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, on_step=None):
        params = list(params)
        defaults = dict(lr=lr, betas=betas, eps=eps)
        super().__init__(params, defaults)
        self._on_step = on_step
        self._flat = None
        self.grad_scale = 1.0
        self._lr_on_device = None

    # ---- flat storage -------------------------------------------------------------------------
    def _params(self):
        return [p for g in self.param_groups for p in g['params']]

    def _flatten(self):
        ps = self._params()
        dev = ps[0].device
        if dev.type != 'cuda':
            raise RuntimeError('FusedAdam: parameters must be CUDA tensors (no CPU path)')
        offs, n = [], 0
        for p in ps:
            offs.append(n)
            n += _align4(p.numel())
        flat_p = torch.zeros(n, device=dev, dtype=torch.float32)
        flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        for p, o in zip(ps, offs):
            v = flat_p[o:o + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
            p.grad = flat_g[o:o + p.numel()].view(p.shape)
        self._flat = dict(p=flat_p, g=flat_g, m=torch.zeros_like(flat_p), v=torch.zeros_like(flat_p), offs=offs, n=n,
                          hyper=torch.zeros(4, device=dev, dtype=torch.float32),
                          step=torch.zeros(1, device=dev, dtype=torch.int32),
                          ptrs=[p.data_ptr() for p in ps])
        self._lr_on_device = None

    def flat(self):
        """(re)build the flat buffers if the parameters moved (e.g. ``module.to()`` or load of new tensors)."""
        ps = self._params()
        if self._flat is None or any(p.data_ptr() != q for p, q in zip(ps, self._flat['ptrs'])):
            self._flatten()
        else:
            # someone may have replaced .grad (zero_grad(set_to_none=True)): re-attach the views
            fg, offs = self._flat['g'], self._flat['offs']
            for p, o in zip(ps, offs):
                if p.grad is None or p.grad.data_ptr() != fg.data_ptr() + o * 4:
                    p.grad = fg[o:o + p.numel()].view(p.shape)
        return self._flat

    def grads_by_param(self):
        return {id(p): p.grad for p in self._params()}

    def zero_flat_grad(self):
        self.flat()['g'].zero_()

    def sync_lr(self):
        """Push param_groups[0]['lr'] to the device copy if it changed (outside any graph capture)."""
        f = self.flat()
        lr = float(self.param_groups[0]['lr'])
        if self._lr_on_device != lr:
            f['hyper'][0:1].copy_(torch.tensor([lr], dtype=torch.float32), non_blocking=False)
            self._lr_on_device = lr

    # ---- the update ---------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None, sync_lr=True):
        loss = closure() if closure is not None else None
        f = self.flat()
        if sync_lr:
            self.sync_lr()
        g = self.param_groups[0]
        b1, b2 = g['betas']
        L.call('pg_adam_step', f['p'].data_ptr(), f['g'].data_ptr(), f['m'].data_ptr(), f['v'].data_ptr(), f['n'],
               f['hyper'].data_ptr(), f['step'].data_ptr(), b1, b2, g['eps'], self.grad_scale, _stream())
        if self._on_step is not None:
            self._on_step()
        return loss

    @torch.no_grad()
    def step_range(self, first, last, bump):
        """Adam update of parameters [first, last) of the flat order only (pg_adam_step_range).  One optimizer step may be
        issued as several ranges on different streams; the caller passes bump=True for exactly one of them, ordered after
        the others, and repacks the operand copies itself (on_step is not called)."""
        f = self.flat()
        offs = f['offs'] + [f['n']]
        lo, hi = offs[first], offs[last]
        g = self.param_groups[0]
        b1, b2 = g['betas']
        L.call('pg_adam_step_range', f['p'].data_ptr() + lo * 4, f['g'].data_ptr() + lo * 4, f['m'].data_ptr() + lo * 4,
               f['v'].data_ptr() + lo * 4, hi - lo, f['hyper'].data_ptr(), f['step'].data_ptr(), b1, b2, g['eps'],
               self.grad_scale, 1 if bump else 0, _stream())

    # ---- checkpointing: the moments and the step count live in the flat device buffers, not in self.state
    def state_dict(self):
        """{'step': int, 'exp_avg': [per-parameter tensors], 'exp_avg_sq': [...], 'lr': float} (CPU tensors, the order of
        the parameters handed to the constructor = module.parameters())."""
        f = self.flat()
        ps = self._params()
        cut = lambda buf: [buf[o:o + p.numel()].view(p.shape).detach().cpu().clone() for p, o in zip(ps, f['offs'])]
        return dict(step=int(f['step'].item()), exp_avg=cut(f['m']), exp_avg_sq=cut(f['v']),
                    lr=float(self.param_groups[0]['lr']))

    def load_state_dict(self, state):
        f = self.flat()
        ps = self._params()
        if len(state['exp_avg']) != len(ps) or any(tuple(a.shape) != tuple(p.shape) for a, p in zip(state['exp_avg'], ps)):
            raise ValueError('FusedAdam.load_state_dict: the saved moments do not match the parameters')
        for name, buf in (('exp_avg', f['m']), ('exp_avg_sq', f['v'])):
            for p, o, a in zip(ps, f['offs'], state[name]):
                buf[o:o + p.numel()].copy_(a.reshape(-1).to(buf.device, torch.float32))
        f['step'].fill_(int(state['step']))
        self.param_groups[0]['lr'] = float(state['lr'])
        self._lr_on_device = None

    def zero_grad(self, set_to_none=False):
        if self._flat is not None:
            self._flat['g'].zero_()
        else:
            super().zero_grad(set_to_none=set_to_none)
