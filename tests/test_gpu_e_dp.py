"""Data-parallel training on two GPUs (NCCL): both ranks end with bit-identical weights, and after the first step they
equal the single-process emulation of SURVEY.md section 8(e): the numpy oracle run per shard from the same initial weights,
gradients averaged, one Adam step.  Needs two GPUs (skipped on a one-GPU box); covers the raw-NCCL single-graph path
(gradient all-reduces captured into the step's CUDA graph and overlapped with the backward pass)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import patchgan_oracle as orc
from tests.golden.cases import CASES

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LR = 1e-3


@pytest.mark.timeout(600)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('raw', ['1', '0'], ids=['raw-nccl-one-graph', 'process-group-split-graphs'])
def test_two_rank_training_matches_the_averaged_gradient_emulation(tmp_path, raw):
    out = str(tmp_path / 'dp')
    env = dict(os.environ, PATCHGAN_B200_RAW_NCCL=raw, MASTER_ADDR='127.0.0.1')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', str(29711 + int(raw)), os.path.join(ROOT, 'tests', 'dp_worker.py'), out, '5']
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=500)
    assert r.returncode == 0, r.stdout[-4000:]
    r0, r1 = np.load(out + '.rank0.npz'), np.load(out + '.rank1.npz')
    assert int(r0['raw_nccl']) == int(raw)
    assert int(r0['graphs']) >= 1                                   # steps 3.. were graph replays
    names = [k for k in r0.files if k.startswith('wN/')]
    for k in names:                                                 # the replicas never drift apart
        assert np.array_equal(r0[k], r1[k]), k
        assert np.array_equal(r0[k.replace('wN/', 'w1/')], r1[k.replace('wN/', 'w1/')]), k
    assert not np.array_equal(r0['wN/' + 'encoder.3.model.DownConv3.weight'], r0['w1/' + 'encoder.3.model.DownConv3.weight'])
    # ---- emulation: per-shard oracle gradients from rank 0's initial weights, averaged, one Adam step
    gk, dk, loss_type, B, _ = CASES['tversky']
    grads = []
    for rank in range(2):
        tr = orc.Trainer(orc.UNet(**gk, seed=11), orc.Discriminator(**dk, seed=12))
        tr.loss_type = loss_type
        x, y = orc.synthetic_batch(B, gk['output_nc'], 256, seed=1234 + rank)
        losses = tr.batch(x, y, train=True)
        grads.append((tr.last['gen_grads'], tr.last['disc_grads']))
        rk = r0 if rank == 0 else r1
        for k, v in losses.items():                                 # per-rank losses of the first step = the reference's per shard
            assert abs(float(rk[f'loss0/{k}']) - v) <= 1e-3 * abs(v), (rank, k)
    g0, d0 = orc.UNet(**gk, seed=11), orc.Discriminator(**dk, seed=12)
    for params, which, prefix in ((g0.params, 0, ''), (d0.params, 1, 'D.')):
        mean = {k: (grads[0][which][k] + grads[1][which][k]) * 0.5 for k in params}
        opt = orc.Adam(params, LR)
        opt.step(mean)
        for k in params:
            diff = np.abs(r0['w1/' + prefix + k] - params[k])
            assert diff.max() <= 2.05 * LR, (k, diff.max())
            assert float(np.mean(diff > 0.5 * LR)) <= 0.05, (k, float(np.mean(diff > 0.5 * LR)))
