"""Loss functions with the reference's names and signatures (/root/reference/patchgan/losses.py:5-39).

These are the user-facing, autograd-differentiable entry points.  ``fc_tversky``, ``MAE_loss`` and
``bce_loss`` run on the single-pass CUDA reducers of libpatchgan_b200 (forward AND gradient); the
``Trainer`` step calls the same kernels directly without going through autograd.
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


def tversky(y_true, y_pred, beta, batch_mean=True):
    """losses.py:5-15.  Not on the Trainer's path (the reference never calls it); kept for API parity and
    implemented with the per-sample sums of the fused reducer."""
    tp, st, sp = _sample_sums(y_true, y_pred)
    fn, fp = st - tp, sp - tp
    tv = tp / (tp + beta * fn + (1. - beta) * fp)
    return torch.mean(1. - tv) if batch_mean else (1. - tv)


def _sample_sums(y_true, y_pred):
    require_cuda(y_pred, 'y_pred')
    B, C, H, W = y_pred.shape
    dev = y_pred.device
    p = _nhwc_f32(y_pred.detach().float())
    t = y_true.detach().float().contiguous()
    part = torch.zeros((B, 8), device=dev)
    L.call('pg_seg_loss_partials', p.data_ptr(), C, t.data_ptr(), None, part.data_ptr(), B, C, H * W, 0, _stream())
    return part[:, 0], part[:, 1], part[:, 2]


def fc_tversky(y_true, y_pred, beta, gamma=0.75, batch_mean=True):
    """losses.py:18-31 (smooth = 1)."""
    if batch_mean:
        return _SegLoss.apply(y_pred, y_true, 'tversky', float(beta), float(gamma))
    tp, st, sp = _sample_sums(y_true, y_pred)
    tv = (tp + 1.) / (tp + beta * (st - tp) + (1. - beta) * (sp - tp) + 1.)
    return torch.pow(1. - tv, gamma)


def MAE_loss(y_true, y_pred):
    """losses.py:34-35: mean |y_true - y_pred| (differentiable wrt y_pred)."""
    return _SegLoss.apply(y_pred, y_true, 'MAE', 0.0, 0.0)


class _BCEConst(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, label):
        require_cuda(p, 'input')
        q = p.detach().float().contiguous()
        n = q.numel()
        losses = torch.zeros(8, device=p.device)
        L.call('pg_bce_const', q.data_ptr(), 1, float(label), 1.0, losses.data_ptr(), 0, None, 0, n, _stream())
        ctx.saved = (q, float(label), p.shape)
        return losses[0].clone()

    @staticmethod
    def backward(ctx, gout):
        q, label, shape = ctx.saved
        # d(mean bce)/dp = (p - t) / max(p (1 - p), 1e-12) / N   (torch's binary_cross_entropy_backward)
        g = (q - label) / torch.clamp(q * (1 - q), min=1e-12) / q.numel()
        return (g * gout).reshape(shape), None


class _BCELoss:
    """nn.BCELoss() stand-in (losses.py:39).  Constant targets (what the Trainer uses: trainer.py:68-69,84,101-102)
    run on the fused reducer; general targets use the weighted-BCE reducer with unit weights."""

    def __call__(self, input, target):
        if target.numel() > 0 and bool((target == target.reshape(-1)[0]).all()):
            return _BCEConst.apply(input, float(target.reshape(-1)[0]))
        if input.dim() != 4:
            input = input.reshape(1, 1, 1, -1)
            target = target.reshape(1, 1, 1, -1)
        B, C, H, W = input.shape
        return _SegLoss.apply(input.reshape(1, 1, B * C * H, W), target.reshape(1, 1, B * C * H, W), 'weighted_bce',
                              0.0, 0.0)


# alias
bce_loss = _BCELoss()
