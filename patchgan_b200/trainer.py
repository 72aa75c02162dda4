"""Trainer with the reference's API (/root/reference/patchgan/trainer.py:16-321) and a fused G+D step.

``Trainer.batch`` computes exactly what the reference's ``batch`` computes (trainer.py:50-115) -- generator
forward, D(fake), segmentation + adversarial generator loss, generator Adam step, D(real), discriminator loss,
discriminator Adam step, the six-entry loss dict -- but schedules it as explicit kernel launches instead of
three autograd graphs:

  * D(fake) and D(real) run as ONE batched discriminator forward over 2B images (InstanceNorm is per-sample, so
    batching cannot change results); the reference's second D(fake) forward (trainer.py:98-99) is bit-identical
    to the first and is not repeated;
  * the generator's backward goes through the discriminator with data-gradients only -- the reference computes
    and then zeroes the discriminator weight-gradients there (trainer.py:89,94);
  * the six ``.item()`` syncs (trainer.py:110-111) become one 4-float device->host copy;
  * with ``torch.distributed`` initialised, each rank steps its own shard and the flat gradient buffers are
    sum-all-reduced over NCCL (1/world folded into the Adam kernel); ``make_optimizers`` broadcasts rank 0's weights so
    the replicas start (and stay) identical.  See ``dp.py`` for how the two all-reduces are scheduled.
"""
import contextlib
import glob
import os
from collections import defaultdict

import warnings

import numpy as np
import torch
import tqdm
from torch.optim.lr_scheduler import ExponentialLR, ReduceLROnPlateau

from . import _lib as L
from . import dp
from . import engine as E
from .engine import _stream, new_act, pack_rows
from .optim import FusedAdam

LOSS_KEYS = ['gen', 'gen_loss', 'gdisc', 'discr', 'discf', 'disc']


def weights_init(net, init_type='normal', scaling=0.02):
    """The reference's ``weights_init`` (trainer.py:327-343) defines an inner function and never applies it, so
    ``module.apply(weights_init)`` leaves the default PyTorch initialisation untouched.  Reproduced as a no-op."""
    return None


_WARNED_NON_FINITE = False


def _warn_non_finite():
    """Forward activations are stored as fp16 by default: |x| > 65504 becomes inf.  InstanceNorm keeps the
    generator's activations O(1); a discriminator without normalisation loaded from a checkpoint with very large weights
    can exceed that range, where the reference (fp32) would not."""
    global _WARNED_NON_FINITE
    if not _WARNED_NON_FINITE:
        _WARNED_NON_FINITE = True
        warnings.warn('patchgan_b200: a loss of this step is not finite.  If the fp32 reference trains this model, an '
                      'activation probably left the fp16 range (65504): set PATCHGAN_B200_FWD_DTYPE=bf16 to store forward '
                      'activations as bfloat16 (fp32 range, 3 fewer mantissa bits).', RuntimeWarning, stacklevel=3)


class PendingLosses:
    """Handle returned by ``Trainer.submit``: the step is queued on the device, ``result()`` waits for its four loss
    scalars and returns the reference's six-entry loss dict (trainer.py:109-113)."""

    def __init__(self, host, event):
        self._host, self._event, self._value = host, event, None

    def done(self):
        return self._value is not None or self._event.query()

    def result(self):
        if self._value is None:
            self._event.synchronize()
            seg, gdisc, discr, discf = (float(v) for v in self._host[:4])
            if not all(np.isfinite(v) for v in (seg, gdisc, discr, discf)):
                _warn_non_finite()
            gen_loss = float(np.float32(seg) + np.float32(gdisc))
            disc_loss = float((np.float32(discf) + np.float32(discr)) / np.float32(2.))
            self._value = dict(zip(LOSS_KEYS, [gen_loss, gen_loss, gdisc, discr, discf, disc_loss]))
            self._host = None
        return self._value


class Trainer:
    '''
        Drives training of a UNet generator against a PatchGAN discriminator
        (same attributes and methods as the reference Trainer).
    '''

    seg_alpha = 200
    loss_type = 'tversky'
    tversky_beta = 0.75
    tversky_gamma = 0.75

    neptune_config = None
    # label ids of the masks (io.py:17,54-56).  When set, batch / submit also take a RAW batch -- x: uint8 (B,3,H,W) image,
    # y: uint8 (B,H,W) label map -- and build the float tensors of io.py:42-56 on the device (patchgan_b200.io.prepare_batch)
    labels = None

    # Replay the step as ONE CUDA graph once a (shape, mode) has been seen GRAPH_WARMUP times (the step is ~140 short
    # launches; the host cannot keep a B200 fed).  PATCHGAN_B200_GRAPH=0 keeps eager launches.
    use_cuda_graph = os.environ.get('PATCHGAN_B200_GRAPH', '1') != '0'
    GRAPH_WARMUP = 2
    # generator layers (first encoder layers) updated at the very end of the step; all others are updated on a side stream
    # while their backward still runs (single GPU, side streams on).  0 = one update at the end.
    LATE_LAYERS = int(os.environ.get('PATCHGAN_B200_LATE_LAYERS', '2'))
    # data-parallel: pieces the generator's early gradient bucket is all-reduced in (Adam / repack of piece i overlap piece i+1)
    DP_BUCKETS = max(1, int(os.environ.get('PATCHGAN_B200_DP_BUCKETS', '1')))    # (2 measured slower at 2 GPUs: 2.59 vs 2.57 ms)

    def __init__(self, generator, discriminator, savefolder, device='cuda'):
        generator.apply(weights_init)
        discriminator.apply(weights_init)
        self.generator = generator
        self.discriminator = discriminator
        self.device = device
        if savefolder[-1] != '/':
            savefolder += '/'
        self.savefolder = savefolder
        if not os.path.exists(savefolder):
            os.mkdir(savefolder)
        self.start = 1
        self._graphs = {}
        self._copy_stream = None      # host->device staging (submit): two input slots, four loss slots
        self._stage = {}
        self._loss_slots = []
        self._submits = 0

    # ------------------------------------------------------------------------------------------
    # one G+D step
    # ------------------------------------------------------------------------------------------
    def make_optimizers(self, gen_lr=1e-3, dsc_lr=1e-3):
        """Adam for both nets exactly as trainer.py:169-172 (lr, betas=(0.9, 0.999)).
        A captured step bakes in the addresses of the optimizers' flat parameter / gradient / moment buffers, so every
        graph captured with the previous optimizers is dropped here (new FusedAdam = new flat buffers, the old ones are
        freed).  Data-parallel: every rank adopts rank 0's weights first -- the reference has no notion of ranks, and
        averaged gradients applied to different initialisations would let the replicas drift apart."""
        self._graphs.clear()
        if dp.world_size() > 1:
            dp.broadcast_parameters(self.generator)
            dp.broadcast_parameters(self.discriminator)
        self.gen_optimizer = FusedAdam(self.generator.parameters(), lr=gen_lr, betas=(0.9, 0.999),
                                       on_step=self.generator._engine().mark_dirty)
        self.disc_optimizer = FusedAdam(self.discriminator.parameters(), lr=dsc_lr, betas=(0.9, 0.999),
                                        on_step=self.discriminator._engine().mark_dirty)
        state = getattr(self, '_resume_optimizer_state', None)
        if state is not None:       # load_last_checkpoint found an optimizer file: continue with its moments / step count
            self.gen_optimizer.load_state_dict(state['generator'])
            self.disc_optimizer.load_state_dict(state['discriminator'])
            self.gen_optimizer.param_groups[0]['lr'] = gen_lr      # (the learning rate is re-derived by train(), as in
            self.disc_optimizer.param_groups[0]['lr'] = dsc_lr     #  the reference: trainer.py:155-157)
            self._resume_optimizer_state = None

    def _graph_signature(self, train):
        """Addresses a captured step depends on beyond its key: the flat optimizer buffers and the parameter storages.
        If anything moved (a second make_optimizers, module.to(), load of new tensors that re-homed the parameters) the
        graph must not be replayed -- it would update freed memory and leave the live weights untouched."""
        sig = [p.data_ptr() for p in self.generator.parameters()] + [p.data_ptr() for p in self.discriminator.parameters()]
        if train:
            for opt in (self.gen_optimizer, self.disc_optimizer):
                f = opt.flat()
                sig += [id(opt), f['p'].data_ptr(), f['g'].data_ptr(), f['m'].data_ptr(), f['v'].data_ptr(),
                        f['hyper'].data_ptr(), f['step'].data_ptr()]
        return tuple(sig)

    def _to_device(self, a):
        if not isinstance(a, torch.Tensor):
            return torch.as_tensor(a, dtype=torch.float).to(self.device)
        return a.to(self.device, non_blocking=True)

    def step_device(self, x, y, train, phase='all'):
        """Issue the whole step on the current stream.  x, y: CUDA float NCHW.  Returns the device tensor
        [seg*alpha, gdisc, discr, discf] (no host sync).
        phase='grads' stops after both backward passes (gradients complete in the flat buffers, all side streams
        joined, no all-reduce, no optimizer step): the data-parallel path replays that part and `update_device` as two
        CUDA graphs with the NCCL all-reduces between them."""
        G, D = self.generator._engine(), self.discriminator._engine()
        gm, dm = self.generator, self.discriminator
        if self.loss_type not in ('tversky', 'weighted_bce', 'MAE'):
            raise ValueError(f'unknown loss_type {self.loss_type!r}')
        B, cin, H, W = x.shape
        cout = gm.output_nc
        if cin != gm.input_nc or y.shape != (B, cout, H, W):
            raise RuntimeError(f'batch: expected x (B,{gm.input_nc},H,W) and y (B,{cout},H,W), got {tuple(x.shape)} '
                               f'and {tuple(y.shape)}')
        if dm.input_nc != cin + cout:
            raise RuntimeError(f'Discriminator.input_nc={dm.input_nc} != {cin}+{cout}')
        dev = x.device
        st = _stream()
        lt = L.LOSS[self.loss_type]
        E.begin_step(dev)
        losses = E.zeros(8, dev)
        world = dp.world_size()
        # Independent chains of the step run on side streams (forked from / joined into the caller's stream):
        #   s_d: D(real) forward during the generator forward, later the whole discriminator update;
        #   s_w: the generator's weight-gradients, off its data-gradient chain;  s_dw: the discriminator's.
        ms = E.Config.streams and L.PROFILER is None
        s_d = s_w = s_dw = s_e = None
        if ms:
            ss = E.side_streams(dev)
            s_d, s_dw, s_e = ss[0], ss[2], ss[6]
            # the generator's weight-gradients: independent launches, spread over NWS streams
            nws = max(1, min(4, int(os.environ.get('PATCHGAN_B200_WSTREAMS', '1'))))
            s_w = ss[1] if nws == 1 else [ss[1]] + ss[3:3 + nws - 1]
        # D(cat(x, y)) forward: under the generator's forward on a side stream (default), or -- DREAL_LATE=1 -- batched with
        # D(cat(x, G(x))) into one 2B-image pass after it (the one-launch conv + InstanceNorm kernels of the generator own
        # every SM while they run, so a concurrent discriminator pass and they only take turns)
        dreal_late = os.environ.get('PATCHGAN_B200_DREAL_LATE', '1') != '0' or bool(getattr(dm, 'batchnorm', False))
        raw_nccl = world > 1 and dp.raw_comm() is not None
        # raw NCCL + side streams: the one-launch conv + norm kernels only leave SMs to NCCL while a collective can actually
        # be in flight beside them (dp.reserve_sms); everywhere else they get the whole GPU
        sm_toggle = (raw_nccl and ms and phase == 'all' and train and self.LATE_LAYERS > 0
                     and os.environ.get('PATCHGAN_B200_DP_SMTOGGLE', '1') != '0' and G.layers_match_parameters()
                     and len(G.specs) > self.LATE_LAYERS)
        if raw_nccl:
            dp.reserve_sms(not sm_toggle)

        def on(stream):
            return torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()

        # ---- inputs.  Both first layers are Conv2d(k4, s2, p1) over 3 / 4 real channels: their inputs are packed
        #      straight into stride-2 im2col matrices (k = c*16 + tap), so the layers are dense pointwise GEMMs.
        #      D input = [fake batch ; real batch] = cat(x, G(x)) ; cat(x, y)  (trainer.py:65, 96).
        im2col = E.taps_enabled() and H % 2 == 0 and W % 2 == 0
        if im2col:
            xs = E.nchw_strides(x)
            xin = E.first_im2col(B, H, W, cin, dev, twin=train)
            E.im2col_fill(xin, 0, x.data_ptr(), xs, cin, 0, B, H, W)
            dboth = E.first_im2col(2 * B, H, W, cin + cout, dev, twin=train)
            L.call('pg_im2col_s2_pair', x.data_ptr(), cin, y.data_ptr(), cout, B, H, W, dboth.ptr,
                   dboth.tw.ptr if dboth.tw is not None else None, dboth.ld, dboth.dt, st)
        else:
            xin = G.pack_input(x, twin=train)
        if im2col:
            pass
        elif D.in_cp in (16, 32):
            dboth = D.new_input(2 * B, H, W, dev, twin=train, zero=False)
            pack_rows(x, None, dboth, 0)          # fake half: [x | (G(x) copied in below) | 0]
            pack_rows(x, y, dboth, B)             # real half: [x | y | 0]
        else:
            dboth = D.new_input(2 * B, H, W, dev, twin=train)
            half = B * H * W * dboth.ld * 2
            for buf in ((dboth, dboth.tw) if dboth.tw is not None else (dboth,)):
                L.call('pg_pack_nchw_f32_to_nhwc_bf16', x.data_ptr(), buf.ptr, B, cin, H, W, buf.ld, 0, buf.dt, st)
                L.call('pg_pack_nchw_f32_to_nhwc_bf16', x.data_ptr(), buf.ptr + half, B, cin, H, W, buf.ld, 0, buf.dt, st)
                L.call('pg_pack_nchw_f32_to_nhwc_bf16', y.data_ptr(), buf.ptr + half, B, cout, H, W, buf.ld, cin, buf.dt,
                       st)

        # ---- D(cat(x, y)) (trainer.py:96-97) does not depend on the generator: side stream, under G's forward
        dctx = D.forward_begin(dboth, save=train)
        if ms and not dreal_late:
            E.fork(s_d)
            with on(s_d):
                D.forward_part(dctx, B, B)

        # ---- generator forward (trainer.py:63) and D(cat(x, G(x))) (trainer.py:65-66)
        if gm.training and gm.use_dropout:
            G.ensure_packed()
            G.bump_seed()
        p, gctx = G.forward(xin, gm.training, save=train)
        if im2col:      # G(x) (f32 NHWC, pixel stride p.ld) -> mask channels of the fake half
            E.im2col_fill(dboth, 0, p.ptr, (H * W * p.ld, 1, W * p.ld, p.ld), cout, cin, B, H, W)
        else:
            L.call('pg_copy_f32_to_bf16_slice', p.ptr, p.ld, dboth.ptr, dboth.ld, cin, cout, B * H * W, dboth.dt, st)
            if dboth.tw is not None:
                L.call('pg_copy_f32_to_bf16_slice', p.ptr, p.ld, dboth.tw.ptr, dboth.ld, cin, cout, B * H * W, 0, st)
        if ms and not dreal_late:
            D.forward_part(dctx, 0, B)
            E.join(s_d)
        elif getattr(dm, 'batchnorm', False):
            # a BatchNorm2d discriminator normalises over the images of one call: D(fake) and D(real) stay separate groups,
            # and the reference's second D(fake) call (trainer.py:98-99) repeats the running-statistics update
            D.forward_part(dctx, 0, B)
            D.forward_part(dctx, B, B)
            D.bn_repeat_update(dctx, 0, B)
        else:
            D.forward_part(dctx, 0, 2 * B)
        pd = dctx[-1][3]
        npatch = B * pd.H * pd.W
        pd_real_ptr = pd.ptr + npatch * pd.ld * 4

        # ---- discriminator losses bce(D(fake),0), bce(D(real),1) (trainer.py:101-103) and update (:105-107):
        #      independent of the generator update (the reference's second D(fake) forward is bit-identical to the first)
        dz = new_act(2 * B, pd.H, pd.W, 16, dev) if train else None
        L.call('pg_bce_const', pd.ptr, pd.ld, 0.0, 0.5, losses.data_ptr(), 3, dz.ptr if train else None, 16, npatch,
               st)
        L.call('pg_bce_const', pd_real_ptr, pd.ld, 1.0, 0.5, losses.data_ptr(), 2,
               dz.ptr + npatch * 16 * 2 if train else None, 16, npatch, st)
        d_work = None
        if train:
            gopt, dopt = self.gen_optimizer, self.disc_optimizer
            gflat, dflat = gopt.flat(), dopt.flat()
            gflat['g'].zero_()
            dflat['g'].zero_()
            dgrads = {n: q.grad for n, q in dm.named_parameters()}
            if ms:
                E.fork(s_d)
            with on(s_d):
                if 'dstep' not in E.SKIP:
                    D.backward(dctx, dz, dgrads, need_dx=False, wstream=s_dw)
                if ms:
                    E.join(s_dw)
                    D.finalize_grads()
                if world > 1 and phase == 'all':
                    # the discriminator's gradients are final: their all-reduce runs on this side stream, underneath the
                    # generator's backward pass (raw NCCL call = a node of the step's graph; else the process group's)
                    if raw_nccl and sm_toggle:
                        pass                         # issued from inside the generator's backward (see `d_exchange` below)
                    elif raw_nccl:
                        dp.raw_all_reduce_sum_(dflat['g'])
                    else:
                        d_work = dp.all_reduce_sum_async(dflat['g'])
                    dopt.grad_scale = 1.0 / world

        # ---- segmentation loss (trainer.py:71-82)
        part = E.zeros((B, 8), dev)
        coef = E.zeros((B, 4), dev)
        chsum = None
        if lt == L.LOSS['weighted_bce']:
            chsum = E.zeros((B, cout), dev)
            L.call('pg_target_chsum', y.data_ptr(), chsum.data_ptr(), B, cout, H * W, st)
        chp = chsum.data_ptr() if chsum is not None else None
        L.call('pg_seg_loss_partials', p.ptr, p.ld, y.data_ptr(), chp, part.data_ptr(), B, cout, H * W, lt, st)
        L.call('pg_seg_loss_finalize', part.data_ptr(), coef.data_ptr(), losses.data_ptr(), 0, B, cout, H * W, lt,
               float(self.tversky_beta), float(self.tversky_gamma), float(self.seg_alpha), st)

        # ---- generator adversarial loss bce(D(fake), 1) (trainer.py:84) and generator update (:87-90)
        dz_g = new_act(B, pd.H, pd.W, 16, dev) if train else None
        L.call('pg_bce_const', pd.ptr, pd.ld, 1.0, 1.0, losses.data_ptr(), 1, dz_g.ptr if train else None, 16,
               npatch, st)
        if train:
            d_dinp = D.backward(dctx, dz_g, None, need_dx=True, nb=B, dx_channels=(cin, cout))
            if ms:
                # the discriminator's Adam step (on s_d) rewrites the weights this data-gradient chain just read
                ev_dread = torch.cuda.Event()
                ev_dread.record()
            # (one output channel: the generator's backward starts with a tap gather of channel 0 only -- trimmed 8-byte rows)
            trim = cout == 1 and E.taps_enabled() and G.packed[-1].taps_ok
            d_raw = new_act(B, H, W, 4 if trim else G.out_cp, dev)
            L.call('pg_gen_out_bwd', p.ptr, p.ld, y.data_ptr(), chp, coef.data_ptr(), d_dinp.ptr, d_dinp.ld, cin,
                   d_raw.ptr, d_raw.ld, B, cout, H * W, lt, L.ACT[gm.final_act], float(self.tversky_beta), st)
            ggrads = {n: q.grad for n, q in gm.named_parameters()}
            K = self.LATE_LAYERS
            if (ms and world == 1 and phase == 'all' and not E.SKIP and K > 0 and len(G.specs) > K
                    and G.layers_match_parameters()):
                # ---- single GPU: the step used to end with a serial tail (last weight-gradient -> finalize -> Adam ->
                #      repack, one kernel on the GPU at a time).  The gradients of every layer but the first K encoder
                #      layers are final ~0.15 ms before the backward pass ends: those layers are finalized, updated and
                #      repacked on s_d underneath the rest of the backward; the tail only handles the K late layers.
                with on(s_d):       # (the discriminator's update first in issue order, the early update queues behind it)
                    s_d.wait_event(ev_dread)
                    dopt.step(sync_lr=False)
                    D.repack()
                nl = len(G.specs)

                KL = K      # (first layer whose gradient is final at the early point; see GeneratorEngine.backward `early`)

                def early_update():
                    ev_m = torch.cuda.Event()
                    for sw in (s_w if isinstance(s_w, list) else [s_w]):   # the weight-gradients of layers KL..
                        if sw in E._FORKED:
                            ev_w = torch.cuda.Event()
                            ev_w.record(sw)
                            s_d.wait_event(ev_w)
                    ev_m.record()                    # the last reads of the operand copies of layers K.. (this stream)
                    s_d.wait_event(ev_m)
                    with on(s_d):
                        G.finalize_grads(partial=True)
                        gopt.step_range(KL, nl, bump=False)
                        G.repack_layers(K, nl, complete=False)     # (layer K-1's copies are still read by its data-gradient)

                G.backward(gctx, d_raw, ggrads, wstream=s_w, early=(K, early_update))
                E.join(s_w)
                G.finalize_grads()
                E.join(s_d)                          # (the early update reads the step count that the next launch bumps)
                gopt.step_range(0, KL, bump=True)
                G.repack_layers(0, K, complete=True)
                E.end_step()
                return losses
            if (ms and raw_nccl and phase == 'all' and K > 0 and len(G.specs) > K and G.layers_match_parameters()):
                # ---- data-parallel, raw NCCL: the generator's gradient buffer is all-reduced in two buckets.  Everything
                #      but the first K encoder layers (99 % of the bytes) is final while the backward of those K layers is
                #      still to run: it goes out on s_d (behind the discriminator's all-reduce, update and repack) and is
                #      followed there by the Adam update and the repack of those layers; the small rest is reduced, updated
                #      and repacked on the main stream at the end.
                #      The main stream WAITS until the bucket's gradients are finalized (ev_fin) before it launches the
                #      remaining one-launch conv + InstanceNorm kernels: those occupy every SM they run on (512 threads x
                #      128 registers), and the finalize kernel and the collective queued behind them used to start only
                #      when the backward pass was over -- no overlap at all (in-graph timeline, tools/timeline.py under
                #      torchrun).  Started first, the collective runs on the SMs pg_set_sm_limit keeps free.
                KL = K      # (first layer whose gradient is final at the early point; see GeneratorEngine.backward `early`)
                off_k = gopt.flat()['offs'][KL]
                nl = len(G.specs)
                gopt.grad_scale = 1.0 / world
                hooks = None
                if sm_toggle:
                    # No collective is in flight during the generator's forward and the first (largest) kernels of its
                    # backward: they were planned for the whole GPU.  The discriminator's all-reduce is released once the
                    # decoder's three big data-gradients are done (ev_big) and runs beside the small layers (64 .. 128 CTAs
                    # whatever the limit); from there on the reservation stays (the main stream never waits for the
                    # collective: with a heavy discriminator -- cfg 5 -- its backward can still be running by then).
                    ev_big = torch.cuda.Event()

                    def d_exchange():
                        ev_big.record()
                        dp.reserve_sms(True)
                        with on(s_d):
                            s_d.wait_event(ev_big)
                            dp.raw_all_reduce_sum_(dflat['g'])
                            s_d.wait_event(ev_dread)
                            dopt.step(sync_lr=False)
                            D.repack()

                    hooks = {('after_dec', 4): d_exchange}
                else:
                    with on(s_d):
                        s_d.wait_event(ev_dread)
                        dopt.step(sync_lr=False)
                        D.repack()
                ev_adam = torch.cuda.Event()

                def early_allreduce():
                    ev_m = torch.cuda.Event()
                    for sw in (s_w if isinstance(s_w, list) else [s_w]):
                        if sw in E._FORKED:
                            ev_w = torch.cuda.Event()
                            ev_w.record(sw)
                            s_d.wait_event(ev_w)
                    ev_m.record()
                    s_d.wait_event(ev_m)
                    ev_fin = torch.cuda.Event()
                    dp.reserve_sms(True)             # the generator's bucket is about to be in flight beside the last layers
                    # the bucket goes out in NB pieces of about equal bytes: Adam and the operand repack of piece i run on
                    # s_e underneath the all-reduce of piece i+1 (the collective runs on the SMs pg_set_sm_limit keeps free)
                    offs = gopt.flat()['offs'] + [gopt.flat()['n']]
                    cuts = [KL]
                    for b in range(1, self.DP_BUCKETS):
                        want = offs[KL] + (offs[nl] - offs[KL]) * b // self.DP_BUCKETS
                        cut = min(range(KL, nl + 1), key=lambda i: abs(offs[i] - want))
                        if cut > cuts[-1] and cut < nl:
                            cuts.append(cut)
                    cuts.append(nl)
                    E.fork(s_e)
                    with on(s_d):
                        G.finalize_grads(partial=True)
                        ev_fin.record()
                        for li, lj in zip(cuts[:-1], cuts[1:]):
                            dp.raw_all_reduce_sum_(gflat['g'], offs[li], offs[lj] - offs[li])
                            ev_b = torch.cuda.Event()
                            ev_b.record()
                            with on(s_e):
                                s_e.wait_event(ev_b)
                                gopt.step_range(li, lj, bump=False)
                                G.repack_layers(max(li, K), lj, complete=False)
                    with on(s_e):
                        ev_adam.record()
                    if os.environ.get('PATCHGAN_B200_DP_HOLD', '1') != '0':
                        torch.cuda.current_stream().wait_event(ev_fin)

                G.backward(gctx, d_raw, ggrads, wstream=s_w, early=(K, early_allreduce), hooks=hooks)
                E.join(s_w)
                G.finalize_grads()
                dp.raw_all_reduce_sum_(gflat['g'], 0, off_k)
                torch.cuda.current_stream().wait_event(ev_adam)      # (the early update reads the step count bumped next)
                gopt.step_range(0, KL, bump=True)
                G.repack_layers(0, K, complete=True)
                E.join(s_d)
                E.join(s_e)
                E.end_step()
                return losses
            if 'gbwd' not in E.SKIP:
                G.backward(gctx, d_raw, ggrads, wstream=s_w)
            if ms:
                E.join(s_w)
                G.finalize_grads()
            if phase == 'grads':
                if ms:
                    E.join(s_d)
                E.end_step()
                return losses
            if world > 1:
                gopt.grad_scale = 1.0 / world
                if raw_nccl:
                    dp.raw_all_reduce_sum_(gflat['g'])
                else:
                    dp.all_reduce_sum_async(gflat['g']).wait()
            if 'adam' not in E.SKIP:
                gopt.step(sync_lr=False)
            with on(s_d):
                if ms:
                    s_d.wait_event(ev_dread)
                if d_work is not None:
                    d_work.wait()
                if 'adam' not in E.SKIP:
                    dopt.step(sync_lr=False)
                    D.repack()        # 16-bit operand copies of the new weights, ready for the next step
            if 'adam' not in E.SKIP:
                G.repack()
            if ms:
                E.join(s_d)
        E.end_step()
        if os.environ.get('PATCHGAN_B200_DEBUGNAN') and not torch.cuda.is_current_stream_capturing():
            torch.cuda.synchronize()
            print('[debugnan] losses', losses.cpu().numpy(), 'train', train, flush=True)
            for i, t in enumerate(E._KEEP):
                tf = t.float()
                nan = int(torch.isnan(tf).sum())
                if nan:
                    print('[debugnan]', i, tuple(t.shape), t.dtype, 'nan', nan, 'of', tf.numel(), flush=True)
        return losses

    def update_device(self):
        """Second half of a data-parallel step: both Adam updates and operand repacks (the flat gradient buffers already
        hold the all-reduced sums; 1/world is folded into the Adam kernel)."""
        G, D = self.generator._engine(), self.discriminator._engine()
        world = dp.world_size()
        self.gen_optimizer.grad_scale = self.disc_optimizer.grad_scale = 1.0 / world
        ms = E.Config.streams and L.PROFILER is None
        dev = next(self.generator.parameters()).device
        s_d = E.side_streams(dev)[0] if ms else None
        if ms:
            E._FORKED.clear()
            E.fork(s_d)
        with (torch.cuda.stream(s_d) if ms else contextlib.nullcontext()):
            self.disc_optimizer.step(sync_lr=False)
            D.repack()
        self.gen_optimizer.step(sync_lr=False)
        G.repack()
        if ms:
            E.join(s_d)

    def _graph_key(self, x, y, train):
        return (tuple(x.shape), tuple(y.shape), bool(train), self.loss_type, float(self.seg_alpha),
                float(self.tversky_beta), float(self.tversky_gamma), self.generator.training,
                self.discriminator.training, x.device.index)

    def step(self, x, y, train):
        """step_device, replayed from a captured CUDA graph when possible.  x, y: CUDA float NCHW (contiguous).
        Returns the device loss tensor of step_device (valid until the next call)."""
        if not self.use_cuda_graph or L.PROFILER is not None or \
                (dp.world_size() > 1 and os.environ.get('PATCHGAN_B200_GRAPH_DP', '1') == '0'):
            return self.step_device(x, y, train)
        # data-parallel: with our own NCCL communicator (dp.raw_comm) the all-reduces are nodes of the step's ONE graph.
        # Without it (PATCHGAN_B200_RAW_NCCL=0, gloo) the process group's collectives stay out of the graphs (capturing them
        # hung on this stack): graph A = everything up to the finished gradients, eager all-reduces, graph B = Adam + repack
        split = dp.world_size() > 1 and dp.raw_comm() is None
        key = self._graph_key(x, y, train)
        ent = self._graphs.get(key)
        if ent is not None and ent['graph'] is not None and ent['sig'] != self._graph_signature(train):
            ent = None                      # buffers moved since the capture: recapture (eager steps meanwhile)
        if ent is None:
            ent = self._graphs[key] = dict(seen=0, graph=None)
        if ent['graph'] is None:
            ent['seen'] += 1
            if ent['seen'] <= self.GRAPH_WARMUP:
                return self.step_device(x, y, train)
            # capture: static input buffers.  The packed operand copies of the weights are brought up to date here
            # (eagerly); inside the graph they are rebuilt right after each optimizer step.
            ent['x'], ent['y'] = x.clone(), y.clone()
            if train:
                self.gen_optimizer.flat()
                self.disc_optimizer.flat()
            self.generator._engine().ensure_packed()
            self.discriminator._engine().ensure_packed()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            # (the process-group watchdog thread may poll events while we capture: check this thread's calls only)
            mode = 'thread_local' if dp.world_size() > 1 else 'global'
            with torch.cuda.graph(g, capture_error_mode=mode):
                ent['losses'] = self.step_device(ent['x'], ent['y'], train, phase='grads' if split else 'all')
            ent['graph'] = g
            ent['sig'] = self._graph_signature(train)
            if split and train:
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, capture_error_mode=mode):
                    self.update_device()
                ent['graph_update'] = g2
            # the capture itself executed nothing: fall through to the first replay
        ent['x'].copy_(x, non_blocking=True)
        ent['y'].copy_(y, non_blocking=True)
        # weights changed from outside (load_state_dict, .to()) since the last step: repack before the replay
        self.generator._engine().ensure_packed()
        self.discriminator._engine().ensure_packed()
        ent['graph'].replay()
        if split and train:
            import torch.distributed as dist
            dist.all_reduce(self.gen_optimizer.flat()['g'])
            dist.all_reduce(self.disc_optimizer.flat()['g'])
            ent['graph_update'].replay()
        return ent['losses']

    def _stage_inputs(self, x, y):
        """Pinned host float tensors: copy into one of two device staging slots on the copy stream, so that the upload of
        batch i+1 runs under the step of batch i (the step's first kernels wait on the slot's event only).  Anything
        else (device tensors, pageable memory, numpy, other dtypes) takes the plain ``.to(device)`` route of
        trainer.py:55-56."""
        def pinned_f32(a):
            return isinstance(a, torch.Tensor) and a.device.type == 'cpu' and a.dtype == torch.float32 and \
                a.is_contiguous() and a.is_pinned()
        if isinstance(x, torch.Tensor) and x.dtype == torch.uint8 and isinstance(y, torch.Tensor) and y.dtype == torch.uint8:
            return self._stage_raw(x, y)
        if not (pinned_f32(x) and pinned_f32(y)) or torch.device(self.device).type != 'cuda':
            xd, yd = self._to_device(x), self._to_device(y)
            if xd.device.type != 'cuda':
                raise RuntimeError('patchgan_b200.Trainer runs on CUDA (sm_100a) only; there is no CPU path')
            return xd.float().contiguous(), yd.float().contiguous(), None
        dev = torch.device(self.device)
        if dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        key = (tuple(x.shape), tuple(y.shape), dev.index)
        slots = self._stage.get(key)
        if slots is None:
            slots = self._stage[key] = [dict(x=torch.empty(x.shape, dtype=torch.float32, device=dev),
                                             y=torch.empty(y.shape, dtype=torch.float32, device=dev),
                                             ready=torch.cuda.Event(), free=None) for _ in range(2)]
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        slot = slots[self._submits % 2]
        cs = self._copy_stream
        if slot['free'] is not None:
            cs.wait_event(slot['free'])          # the step that last read this slot
        with torch.cuda.stream(cs):
            slot['x'].copy_(x, non_blocking=True)
            slot['y'].copy_(y, non_blocking=True)
            slot['ready'].record(cs)
        torch.cuda.current_stream(dev).wait_event(slot['ready'])
        return slot['x'], slot['y'], slot

    def _stage_raw(self, x, y):
        """Raw uint8 batch (image (B,3,H,W), label map (B,H,W)): upload the bytes (copy stream, two staging slots, like
        _stage_inputs) and turn them into the float image / per-label masks on the device."""
        from .io import prepare_batch
        if self.labels is None:
            raise RuntimeError('Trainer.labels (the label ids of the masks) must be set to train from raw uint8 batches')
        dev = torch.device(self.device)
        if dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        nl = len(self.labels)
        B, _, H, W = x.shape
        key = ('u8', tuple(x.shape), nl, dev.index)
        slots = self._stage.get(key)
        if slots is None:
            slots = self._stage[key] = [dict(xu=torch.empty(x.shape, dtype=torch.uint8, device=dev),
                                             yu=torch.empty(y.shape, dtype=torch.uint8, device=dev),
                                             x=torch.empty((B, 3, H, W), dtype=torch.float32, device=dev),
                                             y=torch.empty((B, nl, H, W), dtype=torch.float32, device=dev),
                                             ready=torch.cuda.Event(), free=None) for _ in range(2)]
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        slot = slots[self._submits % 2]
        cs = self._copy_stream
        if slot['free'] is not None:
            cs.wait_event(slot['free'])
        with torch.cuda.stream(cs):
            slot['xu'].copy_(x, non_blocking=True)
            slot['yu'].copy_(y, non_blocking=True)
            prepare_batch(slot['xu'], slot['yu'], self.labels, (H, W), out=(slot['x'], slot['y']))
            slot['ready'].record(cs)
        torch.cuda.current_stream(dev).wait_event(slot['ready'])
        return slot['x'], slot['y'], slot

    def submit(self, x, y, train=False):
        """Queue one ``batch`` without waiting for it: returns a ``PendingLosses`` whose ``result()`` is the loss dict.
        With pinned host inputs the host->device copies go through a copy stream, so a loop that submits batch i+1
        before asking for the result of batch i (``_run_epoch`` does) hides both the upload and the loss read-back.
        Pinned inputs are read asynchronously: leave them unmodified until ``result()`` has returned (a DataLoader with
        ``pin_memory=True`` hands out a fresh tensor per batch)."""
        input_tensor, target_tensor, slot = self._stage_inputs(x, y)
        if train:
            if not hasattr(self, 'gen_optimizer'):
                raise AttributeError("'Trainer' object has no attribute 'gen_optimizer' (call train() or "
                                     "make_optimizers() first)")
            self.gen_optimizer.sync_lr()
            self.disc_optimizer.sync_lr()
        losses = self.step(input_tensor, target_tensor, train)
        main = torch.cuda.current_stream(input_tensor.device)
        if slot is not None:
            if slot['free'] is None:
                slot['free'] = torch.cuda.Event()
            slot['free'].record(main)
        # one device->host copy instead of six .item() calls (trainer.py:110-111); a slot is reused only after its
        # previous owner has been read
        if not self._loss_slots:
            self._loss_slots = [dict(host=torch.empty(8, dtype=torch.float32, pin_memory=True), owner=None)
                                for _ in range(4)]
        ls = self._loss_slots[self._submits % len(self._loss_slots)]
        if ls['owner'] is not None:
            ls['owner'].result()
        ls['host'].copy_(losses, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(main)
        pending = PendingLosses(ls['host'], ev)
        ls['owner'] = pending
        self._submits += 1
        return pending

    def batch(self, x, y, train=False):
        '''
            Train the generator and discriminator on a single batch
        '''
        return self.submit(x, y, train).result()

    # ------------------------------------------------------------------------------------------
    # epoch driver (trainer.py:117-279)
    # ------------------------------------------------------------------------------------------
    def _run_epoch(self, data, train, desc):
        pbar = tqdm.tqdm(data, desc=desc, dynamic_ncols=True)
        if hasattr(data, 'shuffle'):
            data.shuffle()
        sums = defaultdict(float)
        loss_mean = {}
        done = 0

        def account(batch_loss):
            nonlocal done
            done += 1
            for key, value in batch_loss.items():
                sums[key] += value           # O(1) running mean (the reference re-averages a growing list)
                loss_mean[key] = sums[key] / done
            pbar.set_postfix_str(" ".join(f"{key}: {value:.2e}" for key, value in loss_mean.items()))

        # one batch of lookahead: batch i+1 is uploaded and queued before the losses of batch i are read
        pending = None
        for input_img, target_mask in pbar:
            nxt = self.submit(input_img, target_mask, train=train)
            if pending is not None:
                account(pending.result())
            pending = nxt
        if pending is not None:
            account(pending.result())
        return loss_mean

    def train(self, train_data, val_data, epochs, dsc_learning_rate=1.e-3,
              gen_learning_rate=1.e-3, save_freq=10, lr_decay=None, decay_freq=5,
              reduce_on_plateau=False):
        '''
            Training driver: builds the two Adam optimizers, runs `epochs` epochs of `batch(..., train=True)`
            over train_data and `batch(..., train=False)` over val_data, applies LR decay and saves checkpoints.
            Arguments and return value (G_loss_ep, D_loss_ep) are the reference's (trainer.py:117-279).
        '''
        if (lr_decay is not None) and not reduce_on_plateau:
            gen_lr = gen_learning_rate * (lr_decay)**((self.start - 1) / (decay_freq))
            dsc_lr = dsc_learning_rate * (lr_decay)**((self.start - 1) / (decay_freq))
        else:
            gen_lr, dsc_lr = gen_learning_rate, dsc_learning_rate

        cfg = self.neptune_config
        if cfg is not None:
            cfg['model/parameters/gen_learning_rate'] = gen_lr
            cfg['model/parameters/dsc_learning_rate'] = dsc_lr
            cfg['model/parameters/start'] = self.start
            cfg['model/parameters/n_epochs'] = epochs

        self.make_optimizers(gen_lr, dsc_lr)

        if reduce_on_plateau:
            # the reference passes verbose=True, which newer torch rejects, and dereferences neptune_config
            # unconditionally (trainer.py:176-178); both are tolerated here
            gen_scheduler = ReduceLROnPlateau(self.gen_optimizer)
            dsc_scheduler = ReduceLROnPlateau(self.disc_optimizer)
            if cfg is not None:
                cfg['model/parameters/scheduler'] = 'ReduceLROnPlateau'
        elif lr_decay is not None:
            gen_scheduler = ExponentialLR(self.gen_optimizer, gamma=lr_decay)
            dsc_scheduler = ExponentialLR(self.disc_optimizer, gamma=lr_decay)
            if cfg is not None:
                cfg['model/parameters/scheduler'] = 'ExponentialLR'
                cfg['model/parameters/decay_freq'] = decay_freq
                cfg['model/parameters/lr_decay'] = lr_decay
        else:
            gen_scheduler = dsc_scheduler = None

        D_loss_ep, G_loss_ep = [], []
        for epoch in range(self.start, epochs + 1):
            if isinstance(gen_scheduler, ExponentialLR):
                gen_lr = gen_scheduler.get_last_lr()[0]
                dsc_lr = dsc_scheduler.get_last_lr()[0]
            else:
                gen_lr, dsc_lr = gen_learning_rate, dsc_learning_rate
            print(f"Epoch {epoch} -- lr: {gen_lr:5.3e}, {dsc_lr:5.3e}")
            print("-------------------------------------------------------")

            self.generator.train()
            self.discriminator.train()
            loss_mean = self._run_epoch(train_data, True, 'Training: ')
            D_loss_ep.append(loss_mean['disc'])
            G_loss_ep.append(loss_mean['gen'])
            if cfg is not None:
                cfg['train/gen_loss'].append(loss_mean['gen'])
                cfg['train/disc_loss'].append(loss_mean['disc'])

            self.discriminator.eval()
            self.generator.eval()
            loss_mean = self._run_epoch(val_data, False, 'Validation: ')
            if cfg is not None:
                cfg['eval/gen_loss'].append(loss_mean['gen'])
                cfg['eval/disc_loss'].append(loss_mean['disc'])

            if (gen_scheduler is not None) and (dsc_scheduler is not None):
                if isinstance(gen_scheduler, ExponentialLR):
                    if epoch % decay_freq == 0:
                        gen_scheduler.step()
                        dsc_scheduler.step()
                else:
                    # (data-parallel: every rank must see the same metric, or the learning rates diverge)
                    gen_scheduler.step(dp.mean_over_ranks(loss_mean['gen'], self.device))
                    dsc_scheduler.step(dp.mean_over_ranks(loss_mean['disc'], self.device))

            if epoch % save_freq == 0:
                self.save(epoch)

        return G_loss_ep, D_loss_ep

    # ------------------------------------------------------------------------------------------
    # checkpoints (trainer.py:281-321): same file names, fp32 state_dicts in the reference layout
    # ------------------------------------------------------------------------------------------
    def save(self, epoch):
        gen_savefile = f'{self.savefolder}/generator_ep_{epoch:03d}.pth'
        disc_savefile = f'{self.savefolder}/discriminator_ep_{epoch:03d}.pth'
        print(f"Saving to {gen_savefile} and {disc_savefile}")
        if dp.rank() == 0:
            torch.save({k: v.detach().clone() for k, v in self.generator.state_dict().items()}, gen_savefile)
            torch.save({k: v.detach().clone() for k, v in self.discriminator.state_dict().items()}, disc_savefile)
            # not in the reference (it restarts Adam from zero moments on resume, trainer.py:281-287): the optimizer state,
            # in a third file so that the two reference-format files stay exactly what the reference writes
            if hasattr(self, 'gen_optimizer'):
                torch.save(dict(generator=self.gen_optimizer.state_dict(), discriminator=self.disc_optimizer.state_dict()),
                           f'{self.savefolder}/optimizer_ep_{epoch:03d}.pth')

    def load_last_checkpoint(self):
        def epochs_of(prefix):
            files = glob.glob(self.savefolder + f"{prefix}_ep*.pth")
            return {int(os.path.basename(f).replace(f'{prefix}_ep_', '')[:-4]) for f in files}
        gen_epochs, dsc_epochs = epochs_of('generator'), epochs_of('discriminator')
        try:
            assert len(gen_epochs) > 0, "No checkpoints found!"
            start = max(gen_epochs.union(dsc_epochs))
            self.load(f"{self.savefolder}/generator_ep_{start:03d}.pth",
                      f"{self.savefolder}/discriminator_ep_{start:03d}.pth")
            self.start = start + 1
            opt_file = f"{self.savefolder}/optimizer_ep_{start:03d}.pth"
            self._resume_optimizer_state = torch.load(opt_file, map_location='cpu') if os.path.exists(opt_file) else None
        except Exception as e:
            print(e)
            print("Checkpoints not loaded")

    def load(self, generator_save, discriminator_save):
        print(generator_save, discriminator_save)
        dev = next(self.generator.parameters()).device
        self.generator.load_state_dict(torch.load(generator_save, map_location=dev))
        self.discriminator.load_state_dict(torch.load(discriminator_save, map_location=dev))
        gfname = generator_save.split('/')[-1]
        dfname = discriminator_save.split('/')[-1]
        print(f"Loaded checkpoints from {gfname} and {dfname}")
