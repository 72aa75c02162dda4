"""U-Net generator with the reference's constructor, parameter names/shapes and call signature
(/root/reference/patchgan/unet.py:75-134), executed by the sm_100a kernels in ``engine.py``.

The module keeps the reference's sub-module tree (``encoder.{i}.model.DownConv{i}`` /
``decoder.{i}.model.UpConv{i}``) so ``state_dict`` files interchange both ways, but the sub-modules are
parameter holders only: ``forward`` runs the whole network as ONE autograd node (``_UNetFunction``) whose
forward/backward issue the fused kernels.  Tensors must be CUDA tensors; there is no CPU path.
"""
from collections import OrderedDict

import torch
from torch import nn

from . import _lib as L
from . import engine as E
from .engine import Act, GeneratorEngine, _stream, new_act, require_cuda
from .transfer import Transferable

_ACTS = ('tanh', 'relu', 'leakyrelu', 'softmax', 'sigmoid')


class _Holder(nn.Module):
    """Parameter holder reproducing nn.Conv2d / nn.ConvTranspose2d default init (kaiming_uniform(a=sqrt(5)))."""

    def __init__(self, shape, fan_in, bias_n=0):
        super().__init__()
        bound = 1.0 / fan_in ** 0.5
        self.weight = nn.Parameter(torch.empty(shape).uniform_(-bound, bound))
        if bias_n:
            self.bias = nn.Parameter(torch.empty(bias_n).uniform_(-bound, bound))


class DownSampleBlock(nn.Module):
    """unet.py:8-35: Conv2d(k4,s2,p1,bias=False) -> InstanceNorm2d -> act -> [Dropout(0.2)]."""

    def __init__(self, input_filt, output_filt, activation, norm_layer, layer, use_dropout=False, **kwargs):
        super().__init__()
        mods = [(f'DownConv{layer}', _Holder((output_filt, input_filt, 4, 4), input_filt * 16))]
        if norm_layer is nn.BatchNorm2d:      # parameter / buffer holder, same state_dict keys as the reference's module
            mods.append((f'DownNorm{layer}', nn.BatchNorm2d(output_filt)))
        self.model = nn.Sequential(OrderedDict(mods))


class UpSampleBlock(nn.Module):
    """unet.py:38-72: ConvTranspose2d(k4,s2,p1,bias=False) -> [InstanceNorm2d] -> act -> [Dropout(0.2)]."""

    def __init__(self, input_filt, output_filt, activation, norm_layer, layer, batch_norm=True, use_dropout=False,
                 **kwargs):
        super().__init__()
        # torch computes fan_in of a ConvTranspose2d weight (Cin, Cout, 4, 4) from dim 1
        mods = [(f'UpConv{layer}', _Holder((input_filt, output_filt, 4, 4), output_filt * 16))]
        if batch_norm and norm_layer is nn.BatchNorm2d:
            mods.append((f'UpNorm{layer}', nn.BatchNorm2d(output_filt)))
        self.model = nn.Sequential(OrderedDict(mods))


class _UNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, want_hidden, *weights):
        eng = module._engine()
        B, C, H, W = x.shape
        need_grad = any(ctx.needs_input_grad)      # (grad mode is always off inside Function.forward)
        xs = x.contiguous().float()
        if E.taps_enabled() and H % 2 == 0 and W % 2 == 0:
            # first layer as a dense GEMM over the stride-2 im2col of the NCHW input (engine.first_im2col)
            xin = E.first_im2col(B, H, W, C, x.device, twin=need_grad)
            E.im2col_fill(xin, 0, xs.data_ptr(), E.nchw_strides(xs), C, 0, B, H, W)
        else:
            xin = eng.pack_input(xs, twin=need_grad)
        if module.training and module.use_dropout:
            eng.ensure_packed()
            eng.bump_seed()
        p, saved = eng.forward(xin, module.training, save=need_grad)
        out = torch.empty((B, module.output_nc, H, W), device=x.device, dtype=torch.float32)
        L.call('pg_unpack_nhwc_to_nchw_f32', p.ptr, 1, out.data_ptr(), B, module.output_nc, H, W, p.ld, 0, _stream())
        ctx.module, ctx.saved, ctx.p = module, saved, p
        ctx.x_needs_grad = x.requires_grad
        ctx.shape = (B, C, H, W)
        if not want_hidden:
            return out
        # the encoder bottleneck (unet.py:119,131-132), from the SAME forward pass: (B, 8nf, H/128, W/128) NCHW float
        h6 = saved['enc'][6][3] if need_grad else eng.last_bottleneck
        hidden = torch.empty((B, module.nf * 8, h6.H, h6.W), device=x.device, dtype=torch.float32)
        L.call('pg_unpack_nhwc_to_nchw_f32', h6.ptr, h6.dt, hidden.data_ptr(), B, module.nf * 8, h6.H, h6.W, h6.ld, 0,
               _stream())
        ctx.hidden_shape = (h6.H, h6.W, h6.C)
        return out, hidden

    @staticmethod
    def backward(ctx, dout, dhidden=None):
        module, eng = ctx.module, ctx.module._engine()
        B, C, H, W = ctx.shape
        p = ctx.p
        dev = dout.device
        d_hid = None
        if dhidden is not None:
            hh, hw, hc = ctx.hidden_shape
            d_hid = new_act(B, hh, hw, hc, dev, zero=True)
            L.call('pg_pack_nchw_f32_to_nhwc_bf16', dhidden.contiguous().float().data_ptr(), d_hid.ptr, B, module.nf * 8, hh,
                   hw, d_hid.ld, 0, d_hid.dt, _stream())
        # dOut (NCHW float) -> NHWC bf16, then through the final activation
        dpk = new_act(B, H, W, eng.out_cp, dev, zero=True)
        L.call('pg_pack_nchw_f32_to_nhwc_bf16', dout.contiguous().data_ptr(), dpk.ptr, B, module.output_nc, H, W, dpk.ld,
               0, dpk.dt, _stream())
        d_raw = new_act(B, H, W, eng.out_cp, dev)
        L.call('pg_gen_out_bwd', p.ptr, p.ld, None, None, None, dpk.ptr, dpk.ld, 0, d_raw.ptr, d_raw.ld, B,
               module.output_nc, H * W, L.LOSS['none'], L.ACT[module.final_act], 0.0, _stream())
        names = [s.wname for s in eng.specs] + eng.bn_names()
        params = eng.params()
        grads = {n: torch.zeros_like(params[n], dtype=torch.float32) for n in names}
        dx = eng.backward(ctx.saved, d_raw, grads, need_dx=ctx.x_needs_grad, d_hidden=d_hid)
        gx = None
        if ctx.x_needs_grad:
            gx = torch.empty((B, C, H, W), device=dev, dtype=torch.float32)
            L.call('pg_unpack_nhwc_to_nchw_f32', dx.ptr, dx.dt, gx.data_ptr(), B, C, H, W, dx.ld, 0, _stream())
        return (None, gx, None) + tuple(grads[n] for n in names)


class UNet(nn.Module, Transferable):
    def __init__(self, input_nc, output_nc, nf=64, norm_layer=nn.InstanceNorm2d, use_dropout=False,
                 activation='tanh', final_act='softmax'):
        super(UNet, self).__init__()
        if norm_layer not in (nn.InstanceNorm2d, nn.BatchNorm2d):
            raise NotImplementedError('patchgan_b200.UNet implements norm_layer=nn.InstanceNorm2d (the reference default) '
                                      'and nn.BatchNorm2d')
        # BatchNorm2d: batch statistics + affine weight / bias + running buffers; runs the un-fused kernels (conv with the
        # statistics in its epilogue, pg_bn_fold_*, pg_norm_affine_act_*) -- the one-launch conv + norm kernels are
        # InstanceNorm-only
        self.batchnorm = norm_layer is nn.BatchNorm2d
        if activation not in ('tanh', 'relu', 'leakyrelu'):
            raise ValueError(f'activation must be tanh / relu / leakyrelu, got {activation!r}')
        if final_act not in _ACTS:
            raise ValueError(f'final_act must be one of {_ACTS}, got {final_act!r}')
        if output_nc > 16:
            raise NotImplementedError('output_nc > 16 is not supported by the fused loss kernels')
        self.input_nc, self.output_nc, self.nf = input_nc, output_nc, nf
        self.use_dropout, self.activation, self.final_act = use_dropout, activation, final_act

        conv_filts = [nf, nf * 2, nf * 4, nf * 8, nf * 8, nf * 8, nf * 8]
        encoder_layers = []
        prev_filt = input_nc
        for i, filt in enumerate(conv_filts):
            encoder_layers.append(DownSampleBlock(prev_filt, filt, activation, norm_layer, layer=i,
                                                  use_dropout=use_dropout))
            prev_filt = filt
        decoder_layers = []
        for i, filt in enumerate(conv_filts[:-1][::-1]):
            if i == 0:
                decoder_layers.append(UpSampleBlock(prev_filt, filt, activation, norm_layer, layer=i, batch_norm=False))
            else:
                decoder_layers.append(UpSampleBlock(prev_filt * 2, filt, activation, norm_layer, layer=i,
                                                    use_dropout=use_dropout, batch_norm=True))
            prev_filt = filt
        decoder_layers.append(UpSampleBlock(nf * 2, output_nc, final_act, norm_layer, layer=i + 1, batch_norm=False))
        self.encoder = nn.ModuleList(encoder_layers)
        self.decoder = nn.ModuleList(decoder_layers)
        self.__dict__['_eng'] = None

    def _engine(self):
        if self.__dict__.get('_eng') is None:
            self.__dict__['_eng'] = GeneratorEngine(self)
        return self.__dict__['_eng']

    def _weights(self):
        ps = dict(self.named_parameters())
        eng = self._engine()
        return [ps[s.wname] for s in eng.specs] + [ps[n] for n in eng.bn_names()]

    def forward(self, x, return_hidden=False):
        require_cuda(x, 'UNet input')
        if x.dim() != 4 or x.shape[1] != self.input_nc:
            raise RuntimeError(f'UNet expects (B, {self.input_nc}, H, W), got {tuple(x.shape)}')
        # one autograd node; with return_hidden the bottleneck comes from the same pass and is differentiable like the
        # reference's (unet.py:131-134)
        return _UNetFunction.apply(self, x, bool(return_hidden), *self._weights())
