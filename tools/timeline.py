"""In-graph timeline of one training step (cfg3): a %globaltimer stamp after every C-ABI call, on the call's stream, captured
into the step's CUDA graph together with the kernels.  Prints, per call in completion order: stream, end time, and the time
since the previous stamp on the same stream (= the call's duration plus whatever it waited for).
   python tools/timeline.py [out.json]
Data-parallel: run it under torch.distributed.run (N ranks); rank 0 prints its timeline, with a stamp after every raw-NCCL
all-reduce (`ncclAllReduce <MB>`) on the stream it was issued on."""
import ctypes, json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import patchgan_b200 as P
from patchgan_b200 import _lib as L
import bench

from patchgan_b200 import dp

cfg = bench.CONFIGS[os.environ.get('BENCH_CONFIG', 'cfg3')]
rank, world, _local = dp.init_from_env()
dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
torch.cuda.set_device(dev)
torch.manual_seed(0)
G = P.UNet(**cfg['G']).to(dev).train()
D = P.Discriminator(**cfg['D']).to(dev).train()
tr = P.Trainer(G, D, tempfile.mkdtemp(prefix='pgtl'), device=str(dev))
tr.loss_type = cfg['loss_type']
tr.make_optimizers(1e-3, 1e-3)
B, S = cfg['B'], cfg['S']
g = torch.Generator().manual_seed(1234)
x = torch.rand((B, 3, S, S), generator=g).to(dev)
y = (torch.rand((B, 1, S, S), generator=g) > 0.5).float().to(dev)
tr.gen_optimizer.sync_lr(); tr.disc_optimizer.sync_lr()

buf = torch.zeros(4096, dtype=torch.int64, device=dev)
names = []
lib = L.lib()


def stamper(name, args):
    stream = args[-1]
    if not isinstance(stream, ctypes.c_void_p):
        return                      # not a stream call
    sid = stream.value if isinstance(stream, ctypes.c_void_p) else int(stream or 0)
    idx = len(names) % 4096
    d = getattr(args[0], '_obj', None) if args else None
    if isinstance(d, L.ConvDesc):
        name += f" {('conv', 'convT', '1x1')[d.mode]} s{d.stride} B{d.B} {d.Hin}->{d.Hout} C{d.C1}+{d.C2} N{d.N}"
    names.append((name, sid or 0))
    lib.pg_debug_stamp(ctypes.c_void_p(buf.data_ptr() + 8 * idx), ctypes.c_void_p(sid))


_raw_ar = dp.raw_all_reduce_sum_


def stamped_all_reduce(flat, first=0, count=None):
    _raw_ar(flat, first, count)
    if L.STAMPER is not None:
        n = flat.numel() - first if count is None else count
        L.STAMPER(f'ncclAllReduce {n * 4 / 1e6:.1f} MB', [ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)])


dp.raw_all_reduce_sum_ = stamped_all_reduce

for _ in range(tr.GRAPH_WARMUP):
    tr.step(x, y, True)
torch.cuda.synchronize()
L.STAMPER = stamper
names.clear()
tr.step(x, y, True)            # capture (with the stamps) + first replay
L.STAMPER = None
n = len(names)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.zero_()
    tr.step(x, y, True)
torch.cuda.synchronize()
t = buf[:n].cpu().numpy().astype(np.int64)
t0 = t.min()
streams = {}
for (_, sid) in names:
    streams.setdefault(sid, len(streams))
order = np.argsort(t)
last = {}
first_seen = {}
rows = []
for i in range(n):                      # issue order gives the per-stream predecessor
    name, sid = names[i]
    prev = last.get(sid)
    rows.append(dict(i=i, call=name, stream=streams[sid], end_us=(t[i] - t0) / 1e3,
                     since_prev_on_stream_us=None if prev is None else (t[i] - t[prev]) / 1e3))
    last[sid] = i
if rank != 0:
    sys.exit(0)
print(f'{n} calls, span {(t.max() - t0) / 1e3:.1f} us, streams {len(streams)}, world {world}')
for i in order:
    r = rows[i]
    d = r['since_prev_on_stream_us']
    print(f"{r['end_us']:9.1f} us  s{r['stream']}  {'' if d is None else f'{d:8.1f}':>8}  {r['call']}")
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], 'w'))
