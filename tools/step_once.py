"""Run N eager G+D training steps of the bench workload (cfg3) -- the command ncu profiles.
   python tools/step_once.py [steps]   (graphs and side streams off: one kernel at a time, in issue order)"""
import os, sys, tempfile
os.environ['PATCHGAN_B200_GRAPH'] = '0'
os.environ['PATCHGAN_B200_STREAMS'] = '0'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import patchgan_b200 as P
from patchgan_b200 import _lib as L
import bench

cfg = bench.CONFIGS[os.environ.get('BENCH_CONFIG', 'cfg3')]
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device('cuda', 0)
torch.manual_seed(0)
G = P.UNet(**cfg['G']).to(dev).train()
D = P.Discriminator(**cfg['D']).to(dev).train()
tr = P.Trainer(G, D, tempfile.mkdtemp(prefix='pgstep'), device='cuda:0')
tr.loss_type = cfg['loss_type']
tr.make_optimizers(1e-3, 1e-3)
B, S = cfg['B'], cfg['S']
g = torch.Generator().manual_seed(1234)
x = torch.rand((B, 3, S, S), generator=g).to(dev)
y = (torch.rand((B, 1, S, S), generator=g) > 0.5).float().to(dev)
tr.gen_optimizer.sync_lr(); tr.disc_optimizer.sync_lr()
n0 = L.lib().pg_launch_count()
for i in range(steps):
    tr.step_device(x, y, True)
    torch.cuda.synchronize()
    if i == 0:
        print('launches per step', L.lib().pg_launch_count() - n0, flush=True)
print('ok')
