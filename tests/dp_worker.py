"""Worker of tests/test_gpu_e_dp.py: one rank of a data-parallel run (launched by torch.distributed.run, NCCL).
Usage: dp_worker.py OUT_PREFIX [steps]   -> writes OUT_PREFIX.rank{r}.npz"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import patchgan_b200 as P
from oracle import patchgan_oracle as orc            # initial weights + synthetic data only (no oracle arithmetic here)
from patchgan_b200 import dp
from tests.golden.cases import CASES


def main():
    out, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 5
    rank, world, local = dp.init_from_env('nccl')
    dev = torch.device('cuda', local)
    gk, dk, loss_type, B, _ = CASES['tversky']
    # rank r > 0 starts from DIFFERENT weights on purpose: make_optimizers must broadcast rank 0's
    og, od = orc.UNet(**gk, seed=11 + 100 * rank), orc.Discriminator(**dk, seed=12 + 100 * rank)
    G, D = P.UNet(**gk), P.Discriminator(**dk)
    G.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in og.params.items()})
    D.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in od.params.items()})
    tr = P.Trainer(G.to(dev).train(), D.to(dev).train(), tempfile.mkdtemp(prefix=f'pgdp{rank}_'), device=str(dev))
    tr.loss_type = loss_type
    tr.make_optimizers(1e-3, 1e-3)
    x, y = orc.synthetic_batch(B, gk['output_nc'], 256, seed=dp.shard_seed(1234))
    xt, yt = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
    res = {}
    losses = []
    for step in range(steps):
        losses.append(tr.batch(xt, yt, train=True))
        if step == 0:
            for k, p in list(G.named_parameters()) + [('D.' + n, q) for n, q in D.named_parameters()]:
                res['w1/' + k] = p.detach().cpu().numpy().copy()
    for k, p in list(G.named_parameters()) + [('D.' + n, q) for n, q in D.named_parameters()]:
        res['wN/' + k] = p.detach().cpu().numpy().copy()
    for i, l in enumerate(losses):
        for k, v in l.items():
            res[f'loss{i}/{k}'] = np.float64(v)
    res['graphs'] = np.int64(sum(1 for e in tr._graphs.values() if e['graph'] is not None))
    res['raw_nccl'] = np.int64(dp.raw_comm() is not None)
    np.savez(f'{out}.rank{rank}.npz', **res)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
