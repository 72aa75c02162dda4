"""PatchGAN discriminator with the reference's constructor, parameter names/shapes and call signature
(/root/reference/patchgan/disc.py:5-51), executed by the sm_100a kernels in ``engine.py``.

``self.model`` is an ``nn.Sequential`` with parameter holders at the reference's indices (0,2,4,6,8 for
norm=False / n_layers=3; 0,2,5,8,11 for norm=True) and parameter-free placeholders elsewhere, so state_dict
keys match the reference exactly.
"""
import torch
from torch import nn

from . import _lib as L
from . import engine as E
from .engine import DiscriminatorEngine, _stream, new_act, pack_rows, require_cuda
from .transfer import Transferable
from .unet import _Holder


class _Slot(nn.Module):
    """Placeholder for an activation / norm position of the reference Sequential (no parameters)."""

    def __init__(self, what):
        super().__init__()
        self.what = what

    def extra_repr(self):
        return self.what


class _DiscFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        eng = module._engine()
        B, C, H, W = x.shape
        need_grad = any(ctx.needs_input_grad)      # (grad mode is always off inside Function.forward)
        xs = x.contiguous().float()
        if E.taps_enabled() and H % 2 == 0 and W % 2 == 0:
            xin = E.first_im2col(B, H, W, C, x.device, twin=need_grad)
            E.im2col_fill(xin, 0, xs.data_ptr(), E.nchw_strides(xs), C, 0, B, H, W)
        elif eng.in_cp in (16, 32):
            xin = eng.new_input(B, H, W, x.device, twin=need_grad, zero=False)
            pack_rows(xs, None, xin, 0)
        else:
            xin = eng.new_input(B, H, W, x.device, twin=need_grad)
            L.call('pg_pack_nchw_f32_to_nhwc_bf16', xs.data_ptr(), xin.ptr, B, C, H, W, xin.ld, 0, xin.dt, _stream())
            if xin.tw is not None:
                L.call('pg_pack_nchw_f32_to_nhwc_bf16', xs.data_ptr(), xin.tw.ptr, B, C, H, W, xin.ld, 0, 0, _stream())
        p, saved = eng.forward(xin, save=need_grad)
        out = torch.empty((B, 1, p.H, p.W), device=x.device, dtype=torch.float32)
        L.call('pg_unpack_nhwc_to_nchw_f32', p.ptr, 1, out.data_ptr(), B, 1, p.H, p.W, p.ld, 0, _stream())
        ctx.module, ctx.saved, ctx.p = module, saved, p
        ctx.x_needs_grad = x.requires_grad
        ctx.shape = (B, C, H, W)
        return out

    @staticmethod
    def backward(ctx, dout):
        module, eng = ctx.module, ctx.module._engine()
        B, C, H, W = ctx.shape
        p = ctx.p
        dev = dout.device
        dpk = new_act(B, p.H, p.W, 16, dev, zero=True)
        L.call('pg_pack_nchw_f32_to_nhwc_bf16', dout.contiguous().data_ptr(), dpk.ptr, B, 1, p.H, p.W, dpk.ld, 0, dpk.dt,
               _stream())
        d_raw = new_act(B, p.H, p.W, 16, dev)
        # sigmoid backward of the single real channel; channels 1..15 of d_raw are written as zeros
        L.call('pg_gen_out_bwd', p.ptr, p.ld, None, None, None, dpk.ptr, dpk.ld, 0, d_raw.ptr, d_raw.ld, B, 1, p.H * p.W,
               L.LOSS['none'], L.ACT['sigmoid'], 0.0, _stream())
        names = []
        for s in eng.specs:
            names.append(s.wname)
            if s.bias:
                names.append(s.bname)
        names += eng.bn_names()
        params = eng.params()
        grads = {n: torch.zeros_like(params[n], dtype=torch.float32) for n in names}
        dx = eng.backward(ctx.saved, d_raw, grads, need_dx=ctx.x_needs_grad)
        gx = None
        if ctx.x_needs_grad:
            gx = torch.empty((B, C, H, W), device=dev, dtype=torch.float32)
            L.call('pg_unpack_nhwc_to_nchw_f32', dx.ptr, dx.dt, gx.data_ptr(), B, C, H, W, dx.ld, 0, _stream())
        return (None, gx) + tuple(grads[n] for n in names)


class Discriminator(nn.Module, Transferable):
    """Defines a PatchGAN discriminator"""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm=False, norm_layer=nn.InstanceNorm2d):
        super(Discriminator, self).__init__()
        # (norm_layer is only instantiated when norm=True, disc.py:31-32,41-42: with the default norm=False any value is inert)
        if norm and norm_layer not in (nn.InstanceNorm2d, nn.BatchNorm2d):
            raise NotImplementedError('patchgan_b200.Discriminator(norm=True) implements norm_layer=nn.InstanceNorm2d '
                                      '(the reference default) and nn.BatchNorm2d')
        self.input_nc, self.ndf, self.n_layers, self.norm = input_nc, ndf, n_layers, norm
        # BatchNorm2d couples the images of ONE call: the Trainer then runs D(fake) and D(real) as separate groups of the
        # batched pass and repeats the running-statistics update of the reference's second D(fake) call (trainer.py:98-99)
        self.batchnorm = bool(norm) and norm_layer is nn.BatchNorm2d

        def norm_slot(c):
            return nn.BatchNorm2d(c) if self.batchnorm else _Slot('InstanceNorm2d')

        sequence = [_Holder((ndf, input_nc, 4, 4), input_nc * 16, bias_n=ndf), _Slot('LeakyReLU(0.2)')]
        nf_mult = 1
        for n in range(1, n_layers):
            nf_mult_prev, nf_mult = nf_mult, min(2 ** n, 8)
            sequence += [_Holder((ndf * nf_mult, ndf * nf_mult_prev, 4, 4), ndf * nf_mult_prev * 16), _Slot('Tanh')]
            if norm:
                sequence += [norm_slot(ndf * nf_mult)]
        nf_mult_prev, nf_mult = nf_mult, min(2 ** n_layers, 8)
        sequence += [_Holder((ndf * nf_mult, ndf * nf_mult_prev, 4, 4), ndf * nf_mult_prev * 16), _Slot('Tanh')]
        if norm:
            sequence += [norm_slot(ndf * nf_mult)]
        sequence += [_Holder((1, ndf * nf_mult, 4, 4), ndf * nf_mult * 16, bias_n=1), _Slot('Sigmoid')]
        self.model = nn.Sequential(*sequence)
        self.__dict__['_eng'] = None

    def _engine(self):
        if self.__dict__.get('_eng') is None:
            self.__dict__['_eng'] = DiscriminatorEngine(self)
        return self.__dict__['_eng']

    def _params_in_order(self):
        ps = dict(self.named_parameters())
        out = []
        for s in self._engine().specs:
            out.append(ps[s.wname])
            if s.bias:
                out.append(ps[s.bname])
        return out + [ps[n] for n in self._engine().bn_names()]

    def forward(self, input):
        """Standard forward."""
        require_cuda(input, 'Discriminator input')
        if input.dim() != 4 or input.shape[1] != self.input_nc:
            raise RuntimeError(f'Discriminator expects (B, {self.input_nc}, H, W), got {tuple(input.shape)}')
        return _DiscFunction.apply(self, input, *self._params_in_order())
