#!/usr/bin/env python
"""Benchmark of the patchGAN hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py --gpus 1 --steps 20 --warmup 5                 # our CUDA path, one JSON line
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 # CPU arm (oracle port), one JSON line
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                       # data-parallel, weak scaling

A "step" is one full G+D training step (Trainer.batch semantics, trainer.py:50-115) on a synthetic batch of the
BASELINE cfg 3 shape: UNet(3->1, nf=32, leakyrelu, sigmoid) + 3-layer PatchGAN(ndf=64), 256x256, batch 16 per GPU.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# algorithmic minimum GFLOP per image of one G+D step (BASELINE.md / SURVEY.md section 8d)
CONFIGS = {
    'cfg3': dict(workload='cfg3: UNet(3->1,nf=32,leakyrelu,sigmoid)+PatchGAN(ndf=64,L=3) full G+D train step, '
                          '256x256, batch 16/GPU',
                 G=dict(input_nc=3, output_nc=1, nf=32, activation='leakyrelu', final_act='sigmoid', use_dropout=False),
                 D=dict(input_nc=4, ndf=64, n_layers=3), loss_type='tversky', B=16, S=256, gflop_per_img=53.03),
    'cfg4': dict(workload='cfg4: train_coco-shaped UNet(3->7,nf=32,relu,dropout)+PatchGAN(ndf=16,L=5), weighted_bce, '
                          '256x256, batch 16/GPU',
                 G=dict(input_nc=3, output_nc=7, nf=32, activation='relu', final_act='sigmoid', use_dropout=True),
                 D=dict(input_nc=10, ndf=16, n_layers=5), loss_type='weighted_bce', B=16, S=256, gflop_per_img=11.86),
    'cfg5': dict(workload='cfg5: UNet(3->1,nf=64)+PatchGAN(ndf=64,L=4), 1024x1024, batch 4/GPU',
                 G=dict(input_nc=3, output_nc=1, nf=64, activation='leakyrelu', final_act='sigmoid', use_dropout=False),
                 D=dict(input_nc=4, ndf=64, n_layers=4), loss_type='tversky', B=4, S=1024, gflop_per_img=1175.2),
}
CONFIGS['cfg2'] = dict(workload='cfg2: UNet(3->1,nf=32) generator inference-only forward (infer.py path), 256x256, batch 64',
                       G=CONFIGS['cfg3']['G'], D=None, loss_type=None, B=64, S=256, gflop_per_img=3.020, infer=True)
CLOCK_QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
               'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
               'clocks_event_reasons.sw_power_cap')


def ncu_traffic(kernels):
    """DRAM bytes per launch, averaged over the launches of `kernels` (the CUDA kernels behind one C-ABI entry point) in
    the committed ncu launch list of one step (profiles/, tools/summarize_ncu.py); None when no capture is committed."""
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r*_ncu_step_launches.json'))):
        try:
            # (kernel names in the profile keep their template arguments: conv_tc_pers_kernel<0>, conv_res_kernel<1>)
            ks = [k for k in json.load(open(path))['kernels'] if k['kernel'].split('<')[0] in kernels]
            n = sum(k['launches'] for k in ks)
            if n:
                tot = sum(k['dram_MB_per_launch'] * 1e6 * k['launches'] for k in ks)
                best = dict(bytes_per_launch=int(tot / n), launches=n, src=os.path.relpath(path, ROOT))
        except Exception:
            pass
    return best


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(src='measured', hbm=p['hbm_gbs'], tf_burst=p['bf16_tflops'], tf_sustained=p['bf16_tflops_sustained'])
    return dict(src='fallback', hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0)


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference step, all host threads, bounded sample
# ----------------------------------------------------------------------------------------------------------------
CPU_SAMPLE = ("oracle/torch_port.py: the reference step (trainer.py:50-115, unet.py, disc.py, optim.Adam) restated on "
              "torch CPU operators -- the reference's own arithmetic library -- all host threads, fp32")


def cpu_step_rate(cfg, steps, warmup, batch=4):
    import numpy as np
    import torch
    from oracle import patchgan_oracle as orc
    from oracle import torch_port as tp
    torch.set_num_threads(os.cpu_count())
    gk = cfg['G']
    og = orc.UNet(gk['input_nc'], gk['output_nc'], gk['nf'], use_dropout=False, activation=gk['activation'],
                  final_act=gk['final_act'], seed=0)
    od = orc.Discriminator(cfg['D']['input_nc'], cfg['D']['ndf'], cfg['D']['n_layers'], seed=1)
    st = tp.Step(og.params, od.params, gk, cfg['D'], cfg['loss_type'])
    x, y = orc.synthetic_batch(batch, gk['output_nc'], cfg['S'], seed=1234)
    x, y = torch.from_numpy(x), torch.from_numpy(y)
    for _ in range(warmup):
        st.batch(x, y, train=True)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        st.batch(x, y, train=True)
        ts.append(time.perf_counter() - t0)
    sec = float(np.median(ts))
    return batch / sec, sec


def cpu_batch(cfg):
    """Batch of one CPU step: the configuration's own per-GPU batch at 256 x 256 (a step is well under a second per image on
    the host), one image at 1024 x 1024 (bounded sample; the workload string of that line says so)."""
    return cfg['B'] if cfg['S'] <= 256 else 1


def run_reference(args, cfg):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    if cfg.get('infer'):
        print(json.dumps(dict(impl='reference', unavailable='the CPU arm times the training step (cfg3/4/5) only')), flush=True)
        return
    steps, warmup = max(1, min(args.steps, 10)), max(1, min(args.warmup, 2))
    batch = cpu_batch(cfg)
    if cfg['S'] > 256:
        steps, warmup = min(steps, 3), 1          # a 1024 x 1024 step takes tens of seconds on the host
    rate, sec = cpu_step_rate(cfg, steps, warmup, batch)
    cores = os.cpu_count()
    sample = f'{CPU_SAMPLE}; {warmup} warm-up + {steps} timed step(s) of batch {batch} at {cfg["S"]}x{cfg["S"]}, median'
    line = dict(metric='train_img_per_s', value=round(rate, 3), unit='img/s', n_gpus=args.gpus, steps=steps, warmup=warmup,
                ms_per_step=round(sec * 1e3, 2), higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                data='synthetic', impl='reference',
                config=dict(workload=cfg['workload'] + ('' if batch == cfg['B'] else f' [CPU sample: batch {batch}]'),
                            cpu_batch=batch, same_config=batch == cfg['B']),
                cpu_baseline=dict(value=round(rate, 3), unit='img/s', cores=cores, kind='port', sample=sample),
                e2e=dict(value=round(rate, 3), unit='img/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
class CallProfiler:
    """Times every C-ABI call with CUDA events on the launching stream (used in a separate pass after the timed
    region so the headline number is not perturbed)."""

    def __init__(self, torch, lib):
        self.torch, self.lib, self.rec, self._flops, self._tag = torch, lib, [], 0.0, ''

    def note(self, flops, tag=''):
        self._flops = flops
        self._tag = tag

    def begin(self, name):
        e = self.torch.cuda.Event(enable_timing=True)
        e.record()
        fl, self._flops = self._flops, 0.0
        tag, self._tag = self._tag, ''
        return (name, fl, e, tag)

    def end(self, tok):
        e = self.torch.cuda.Event(enable_timing=True)
        e.record()
        # same CUDA kernels behind three entry points each: + fused statistics, + TMA bulk-reduce epilogue
        name = {'pg_conv_fwd_stats': 'pg_conv_fwd', 'pg_conv_dgrad_act': 'pg_conv_fwd',
                'pg_conv_wgrad_tapmajor': 'pg_conv_wgrad'}.get(tok[0], tok[0])
        if name in ('pg_conv_fwd', 'pg_conv_wgrad'):
            name += {2: ':tcgen05', 3: ':skinny'}.get(self.lib.pg_last_conv_impl(), ':simt')
        self.rec.append((name, tok[1], tok[2], e, tok[3]))

    def summary(self):
        self.torch.cuda.synchronize()
        agg = {}
        self.detail = []
        for name, fl, e0, e1, tag in self.rec:
            ms = e0.elapsed_time(e1)
            a = agg.setdefault(name, [0, 0.0, 0.0])
            a[0] += 1
            a[1] += ms
            a[2] += fl
            self.detail.append(dict(call=name, shape=tag, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1) if fl else None))
        return agg


def sample_clocks(dev_index):
    try:
        return subprocess.Popen(['nvidia-smi', f'--id={dev_index}', f'--query-gpu={CLOCK_QUERY}', '--format=csv,noheader,nounits',
                                 '-lms', '50'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return None


def finish_clocks(proc):
    if proc is None:
        return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
    proc.terminate()
    try:
        out, _ = proc.communicate(timeout=5)
    except Exception:
        proc.kill()
        out = ''
    sm, mx, reasons = [], [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for ln in out.strip().splitlines():
        f = [t.strip() for t in ln.split(',')]
        if len(f) < 9:
            continue
        try:
            sm.append(float(f[1]))
            mx.append(float(f[2]))
        except ValueError:
            continue
        for n, v in zip(names, f[5:9]):
            if v.lower().startswith('active'):
                reasons.add(n)
    sm.sort()
    return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                samples=len(sm))


def run_infer(args, cfg):
    print(json.dumps(infer_numbers(args, cfg, full=True)), flush=True)


def infer_numbers(args, cfg, full=False):
    """cfg2: generator forward only, through the module call a user makes (infer.py:155-170: eval(), no_grad).
    full=False: the compact form embedded as `infer` in the training line."""
    import torch
    import patchgan_b200 as P
    from patchgan_b200 import _lib as L
    from patchgan_b200.engine import Config
    sys.stdout.flush()
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    B, S = cfg['B'], cfg['S']
    torch.manual_seed(0)
    G = P.UNet(**cfg['G']).to(dev).eval()
    gen = torch.Generator().manual_seed(1234)
    x_host = torch.rand((B, 3, S, S), generator=gen).pin_memory()
    x_dev = x_host.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out_host = torch.empty((B, cfg['G']['output_nc'], S, S), dtype=torch.float32).pin_memory()
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            G(x_dev)
        torch.cuda.synchronize()
        clocks = sample_clocks(0)
        n0 = L.lib().pg_launch_count()
        evs = []
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); G(x_dev); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        launches = (L.lib().pg_launch_count() - n0) // args.steps
        ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out_host.copy_(G(x_host.to(dev, non_blocking=True)), non_blocking=True)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / args.steps
    pk = peaks()
    img_s = B / (ms * 1e-3)
    tf = img_s * cfg['gflop_per_img'] / 1e3
    line = dict(metric='infer_img_per_s', value=round(img_s, 2), unit='img/s', n_gpus=1, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=round(ms, 4), higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype=('f16' if Config.fwd_dt == L.DT_F16 else 'bf16') + ' operands, f32 accumulate', data='synthetic',
                config=dict(workload=cfg['workload'], global_batch=B, image=S, l2='flushed (256 MB write) between timed steps',
                            cuda_graph=False),
                clocks=finish_clocks(clocks),
                e2e=dict(value=round(B / e2e_s, 2), unit='img/s', h2d_bytes_per_step=int(x_host.numel() * 4),
                         d2h_bytes_per_step=int(out_host.numel() * 4), ms_per_step=round(e2e_s * 1e3, 4)),
                gpu_launches=int(launches),
                roofline=dict(bound='tensor', kernel='generator forward (all launches)', achieved=round(tf, 2),
                              peak=pk['tf_sustained'], unit='TFLOP/s', frac=round(tf / pk['tf_sustained'], 4), traffic=None))
    if full:
        return line
    return dict(metric=line['metric'], workload=cfg['workload'], value=line['value'], unit='img/s', steps=line['steps'],
                ms_per_step=line['ms_per_step'], e2e=line['e2e'], gpu_launches=line['gpu_launches'],
                roofline_frac=line['roofline']['frac'])


def run_ours(args, cfg):
    import torch
    import patchgan_b200 as P
    from patchgan_b200 import _lib as L
    from patchgan_b200 import dp
    from patchgan_b200.engine import Config

    # keep stdout clean for the single JSON line: libraries (NCCL banner, ...) write to fd 1 during the run
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    rank, world, local = dp.init_from_env('nccl')
    if world != args.gpus and world > 1:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}')
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    lib = L.lib()
    B, S = cfg['B'], cfg['S']

    torch.manual_seed(0)
    G = P.UNet(**cfg['G']).to(dev).train()
    D = P.Discriminator(**cfg['D']).to(dev).train()
    dp.broadcast_parameters(G)
    dp.broadcast_parameters(D)
    tr = P.Trainer(G, D, tempfile.mkdtemp(prefix='pgbench'), device=str(dev))
    tr.loss_type = cfg['loss_type']
    tr.make_optimizers(1e-3, 1e-3)

    # synthetic RAW batch, the form a decoded COCO-stuff sample has (io.py:42-43): uint8 RGB image + uint8 label map; the
    # device pipeline (patchgan_b200/io.py, pg_prep_batch_u8) turns it into the float image and the per-label masks
    from patchgan_b200.io import prepare_batch
    gen = torch.Generator().manual_seed(dp.shard_seed(1234))
    cout = cfg['G']['output_nc']
    x_host = torch.randint(0, 256, (B, 3, S, S), generator=gen, dtype=torch.uint8).pin_memory()
    y_host = torch.randint(0, cout + 1, (B, S, S), generator=gen, dtype=torch.uint8).pin_memory()      # raw label map
    tr.labels = list(range(2, cout + 2))         # mask i = (label map + 1 == i + 2) = (label map == i + 1), some pixels unlabelled
    x_dev, y_dev = prepare_batch(x_host.to(dev), y_host.to(dev), tr.labels, (S, S))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    tr.gen_optimizer.sync_lr()
    tr.disc_optimizer.sync_lr()
    for _ in range(max(args.warmup, 3) + 1):     # (includes the CUDA-graph capture of the step)
        tr.step(x_dev, y_dev, True)
    barrier()

    clocks = sample_clocks(local) if rank == 0 else None
    # ---- device-timed region: inputs resident in HBM, CUDA events on the launching stream, L2 flushed between steps
    evs = []
    barrier()
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tr.step(x_dev, y_dev, True)
        e1.record()
        evs.append((e0, e1))
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    dev_ms = dp.max_over_ranks(dev_ms, dev)

    # ---- end to end through the public API: pinned host inputs -> Trainer.batch -> loss dict on the host
    for _ in range(3):          # untimed: staging slots, pinned loss slots and the copy stream are created on first use
        tr.batch(x_host, y_host, train=True)
    #      (a) Trainer.batch: upload, step, loss read-back, strictly one after the other (the reference's calling pattern)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        losses = tr.batch(x_host, y_host, train=True)
    torch.cuda.synchronize()
    e2e_sync_s = dp.max_over_ranks(time.perf_counter() - t0, dev)
    #      (b) Trainer.submit / .result() with one batch of lookahead -- the loop Trainer.train runs: every step still
    #      uploads its own inputs from pinned host memory and reads its own losses back, all inside the timed region
    barrier()
    t0 = time.perf_counter()
    pending = None
    for _ in range(args.steps):
        nxt = tr.submit(x_host, y_host, train=True)
        if pending is not None:
            losses = pending.result()
        pending = nxt
    losses = pending.result()
    torch.cuda.synchronize()
    e2e_s = dp.max_over_ranks(time.perf_counter() - t0, dev)
    clock_info = finish_clocks(clocks) if rank == 0 else None

    # ---- per-kernel pass (separate from the timed region): CUDA events around every C-ABI call
    prof = CallProfiler(torch, lib)
    L.PROFILER = prof
    nprof = 2
    n0 = lib.pg_launch_count()
    for _ in range(nprof):
        tr.step_device(x_dev, y_dev, True)
    launches = (lib.pg_launch_count() - n0) // nprof      # kernels of this library per step (graph replays launch the same)
    L.PROFILER = None
    fallbacks = lib.pg_fallback_count()
    if fallbacks:
        raise SystemExit(f'{fallbacks} convolution call(s) of this run had no tcgen05 plan and ran on the CUDA-core kernel')
    agg = prof.summary()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    if rank != 0:
        return
    if args.detail:
        with open(args.detail, 'w') as f:
            json.dump(prof.detail[len(prof.detail) // nprof:], f, indent=0)
    pk = peaks()
    total_prof_ms = sum(a[1] for a in agg.values())
    top = sorted(agg.items(), key=lambda kv: -kv[1][1])
    shares = {k: round(v[1] / total_prof_ms, 4) for k, v in top[:8]}
    conv = [(k, v) for k, v in top if v[2] > 0]
    kname, (cnt, ms, fl) = conv[0]
    achieved_tf = fl / (ms * 1e-3) / 1e12
    cuda_kernel = {'pg_conv_fwd:tcgen05': ['conv_tc_pers_kernel', 'conv_tc_kernel'],
                   'pg_conv_wgrad:tcgen05': ['wgrad_tc_kernel'], 'pg_conv_wgrad_group': ['wgrad_group_kernel'],
                   'pg_conv_norm_fwd': ['conv_res_kernel'], 'pg_conv_dgrad_norm_bwd': ['conv_res_kernel']}.get(kname)
    tr_info = ncu_traffic(cuda_kernel) if cuda_kernel else None
    roof = dict(bound='tensor', kernel=kname, cuda_kernel=cuda_kernel, achieved=round(achieved_tf, 2), peak=pk['tf_sustained'],
                peak_src=pk['src'] + ' (sustained cuBLAS bf16)', unit='TFLOP/s', frac=round(achieved_tf / pk['tf_sustained'], 4),
                traffic=tr_info['bytes_per_launch'] if tr_info else None,
                traffic_src=tr_info['src'] if tr_info else None,
                launches_per_step=cnt // nprof, ms_per_step=round(ms / nprof, 4), flops_per_step=fl / nprof,
                flops_per_launch=fl / max(cnt, 1),
                how='CUDA events around every launch of this kernel in a separate eager pass (side streams and graph off); '
                    'achieved = sum of 2*MACs as launched / sum of durations')
    img_s = world * B * args.steps / (dev_ms * 1e-3)
    step_tf = img_s * cfg['gflop_per_img'] / 1e3 / world          # per GPU: the peak below is one GPU's
    line = dict(metric='train_img_per_s', value=round(img_s, 2), unit='img/s', n_gpus=world, steps=args.steps,
                warmup=max(args.warmup, 3), ms_per_step=round(dev_ms / args.steps, 4), higher_is_better=True,
                scaling='weak', vs_baseline=None,
                dtype=('f16' if Config.fwd_dt == L.DT_F16 else 'bf16') + ' forward operands, bf16 gradients, f32 accumulate',
                data='synthetic',
                config=dict(workload=cfg['workload'], global_batch=world * B, per_gpu_batch=B, image=S,
                            parallelism=f'dp{world}', l2='flushed (256 MB write) between timed steps',
                            conv_impl={0: 'auto', 1: 'simt', 2: 'tcgen05'}[Config.impl],
                            cuda_graph=bool(tr.use_cuda_graph), side_streams=bool(Config.streams)),
                clocks=clock_info,
                e2e=dict(value=round(world * B * args.steps / e2e_s, 2), unit='img/s',
                         h2d_bytes_per_step=int(x_host.numel() + y_host.numel()), d2h_bytes_per_step=32,
                         ms_per_step=round(e2e_s / args.steps * 1e3, 4),
                         api='Trainer.submit(x_u8_host, labelmap_u8_host).result(): raw uint8 upload, float image + masks built on the '
                             'device (pg_prep_batch_u8), one batch of lookahead (the loop of Trainer.train)',
                         sync_batch_value=round(world * B * args.steps / e2e_sync_s, 2),
                         sync_batch_api='Trainer.batch(x_u8_host, labelmap_u8_host): upload, step, read-back in sequence'),
                gpu_launches=int(launches), roofline=roof,
                step_tensor=dict(algorithmic_gflop_per_img=cfg['gflop_per_img'], achieved_tflops_per_gpu=round(step_tf, 2),
                                 frac_of_peak=round(step_tf / pk['tf_sustained'], 4)),
                cuda_core_fallbacks=int(fallbacks),
                kernel_time_shares=shares, last_losses={k: round(v, 5) for k, v in losses.items()})
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_batch(cfg)
        nst = 5 if S <= 256 else 2
        rate, sec = cpu_step_rate(cfg, nst, 1, cb)
        line['cpu_baseline'] = dict(value=round(rate, 3), unit='img/s', cores=os.cpu_count(), kind='port',
                                    sample=f'{CPU_SAMPLE}; 1 warm-up + {nst} timed steps of batch {cb}, median')
    if world == 1 and not args.no_infer and args.config == 'cfg3':
        # BASELINE.json configs[1] (generator inference forward, batch 64) inside the same driver-run line
        line['infer'] = infer_numbers(args, CONFIGS['cfg2'])
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='cfg3', choices=list(CONFIGS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-infer', action='store_true', help='skip the cfg2 inference numbers added to the cfg3 line')
    ap.add_argument('--detail', default=None, help='write the per-launch timing of one profiled step to this JSON file')
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == 'reference':
        run_reference(args, cfg)
    elif cfg.get('infer'):
        run_infer(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == '__main__':
    main()
