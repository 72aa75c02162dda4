"""Loss functions with the reference's names and signatures (/root/reference/patchgan/losses.py:5-39).

These are the user-facing, autograd-differentiable entry points (every function, every mode, like the reference's).
They run on the single-pass CUDA reducers of libpatchgan_b200 (forward AND gradient); the ``Trainer`` step calls the
fused kernels directly without going through autograd.
"""
import ctypes

import torch

from . import _lib as L
from .engine import _stream, require_cuda


def _nhwc_f32(p):
    """NCHW float -> contiguous NHWC float view-copy (layout plumbing for the standalone loss API)."""
    return p.permute(0, 2, 3, 1).contiguous()


class _SegLoss(torch.autograd.Function):
    """loss_type in {'tversky','weighted_bce','MAE'}; returns the UNSCALED loss (seg_alpha = 1)."""

    @staticmethod
    def forward(ctx, y_pred, y_true, loss_type, beta, gamma):
        require_cuda(y_pred, 'y_pred')
        require_cuda(y_true, 'y_true')
        B, C, H, W = y_pred.shape
        if C > 16:
            raise NotImplementedError('more than 16 output channels')
        dev = y_pred.device
        p = _nhwc_f32(y_pred.detach().float())
        t = y_true.detach().float().contiguous()
        lt = L.LOSS[loss_type]
        st = _stream()
        part = torch.zeros((B, 8), device=dev)
        coef = torch.zeros((B, 4), device=dev)
        losses = torch.zeros(8, device=dev)
        chsum = None
        if lt == L.LOSS['weighted_bce']:
            chsum = torch.zeros((B, C), device=dev)
            L.call('pg_target_chsum', t.data_ptr(), chsum.data_ptr(), B, C, H * W, st)
        L.call('pg_seg_loss_partials', p.data_ptr(), C, t.data_ptr(), chsum.data_ptr() if chsum is not None else None,
               part.data_ptr(), B, C, H * W, lt, st)
        L.call('pg_seg_loss_finalize', part.data_ptr(), coef.data_ptr(), losses.data_ptr(), 0, B, C, H * W, lt, beta,
               gamma, 1.0, st)
        ctx.saved = (p, t, chsum, coef, lt, beta, (B, C, H, W))
        return losses[0].clone()

    @staticmethod
    def backward(ctx, gout):
        p, t, chsum, coef, lt, beta, (B, C, H, W) = ctx.saved
        dev = p.device
        ld = max(8, (C + 7) // 8 * 8)
        dx = torch.empty((B, H, W, ld), device=dev, dtype=torch.bfloat16)
        # final_act = none: plain d(loss)/d(p); result in bf16 NHWC, then back to NCHW float
        L.call('pg_gen_out_bwd', p.data_ptr(), C, t.data_ptr(), chsum.data_ptr() if chsum is not None else None,
               coef.data_ptr(), None, 0, 0, dx.data_ptr(), ld, B, C, H * W, lt, 0, beta, _stream())
        g = torch.empty((B, C, H, W), device=dev, dtype=torch.float32)
        L.call('pg_unpack_nhwc_to_nchw_f32', dx.data_ptr(), 0, g.data_ptr(), B, C, H, W, ld, 0, _stream())
        return g * gout, None, None, None, None


class _SampleSums(torch.autograd.Function):
    """Per-sample (sum t*p, sum t, sum p) over (C, H, W) from the fused reducer, differentiable wrt the prediction
    (d tp/dp = t, d sp/dp = 1): what tversky / fc_tversky(batch_mean=False) are made of (losses.py:6-8, 20-22)."""

    @staticmethod
    def forward(ctx, y_pred, y_true):
        require_cuda(y_pred, 'y_pred')
        require_cuda(y_true, 'y_true')
        B, C, H, W = y_pred.shape
        if C > 16:
            raise NotImplementedError('more than 16 output channels')
        p = _nhwc_f32(y_pred.detach().float())
        t = y_true.detach().float().contiguous()
        part = torch.zeros((B, 8), device=y_pred.device)
        L.call('pg_seg_loss_partials', p.data_ptr(), C, t.data_ptr(), None, part.data_ptr(), B, C, H * W, 0, _stream())
        ctx.saved = (t, (B, C, H, W))
        tp, st, sp = part[:, 0].clone(), part[:, 1].clone(), part[:, 2].clone()
        ctx.mark_non_differentiable(st)        # sum t does not depend on the prediction
        return tp, st, sp

    @staticmethod
    def backward(ctx, g_tp, g_st, g_sp):
        t, (B, C, H, W) = ctx.saved
        dp = torch.empty((B, C, H, W), device=t.device, dtype=torch.float32)
        L.call('pg_sample_sums_bwd', t.data_ptr(), g_tp.contiguous().float().data_ptr(), g_sp.contiguous().float().data_ptr(),
               dp.data_ptr(), B, C * H * W, _stream())
        return dp, None


def tversky(y_true, y_pred, beta, batch_mean=True):
    """losses.py:5-15.  Not on the Trainer's path (the reference never calls it); the three reductions are one pass of the
    fused reducer, the per-sample arithmetic on the B results is left to torch (differentiable like the reference)."""
    tp, st, sp = _SampleSums.apply(y_pred, y_true)
    fn, fp = st - tp, sp - tp
    tv = tp / (tp + beta * fn + (1. - beta) * fp)
    return torch.mean(1. - tv) if batch_mean else (1. - tv)


def fc_tversky(y_true, y_pred, beta, gamma=0.75, batch_mean=True):
    """losses.py:18-31 (smooth = 1)."""
    if batch_mean:
        return _SegLoss.apply(y_pred, y_true, 'tversky', float(beta), float(gamma))
    tp, st, sp = _SampleSums.apply(y_pred, y_true)
    tv = (tp + 1.) / (tp + beta * (st - tp) + (1. - beta) * (sp - tp) + 1.)
    return torch.pow(1. - tv, gamma)


def MAE_loss(y_true, y_pred):
    """losses.py:34-35: mean |y_true - y_pred| (differentiable wrt y_pred)."""
    return _SegLoss.apply(y_pred, y_true, 'MAE', 0.0, 0.0)


class _BCEMean(torch.autograd.Function):
    """nn.BCELoss()(input, target) (losses.py:39): mean over all elements, log clamped at -100; forward and gradient are
    one kernel each, no host synchronisation (the upstream gradient is read on the device)."""

    @staticmethod
    def forward(ctx, p, t):
        require_cuda(p, 'input')
        require_cuda(t, 'target')
        if p.shape != t.shape:
            raise ValueError(f'Using a target size ({tuple(t.shape)}) that is different to the input size '
                             f'({tuple(p.shape)}) is deprecated. Please ensure they have the same size.')
        q = p.detach().float().contiguous()
        tt = t.detach().float().contiguous()
        losses = torch.zeros(8, device=p.device)
        L.call('pg_bce_mean', q.data_ptr(), tt.data_ptr(), q.numel(), losses.data_ptr(), 0, _stream())
        ctx.saved = (q, tt, p.shape)
        return losses[0].clone()

    @staticmethod
    def backward(ctx, gout):
        q, tt, shape = ctx.saved
        dp = torch.empty_like(q)
        L.call('pg_bce_mean_bwd', q.data_ptr(), tt.data_ptr(), q.numel(), gout.contiguous().float().data_ptr(),
               dp.data_ptr(), _stream())
        return dp.reshape(shape), None


class _BCELoss:
    """nn.BCELoss() stand-in (losses.py:39)."""

    def __call__(self, input, target):
        return _BCEMean.apply(input, target)


# alias
bce_loss = _BCELoss()
