"""CPU tests of the data-parallel plumbing (world_size 2, gloo): the helpers in patchgan_b200/dp.py and the DP
semantics the Trainer implements on NCCL -- each rank runs the reference step on its own shard, the flat gradient
buffers are sum-all-reduced and the 1/world factor is applied inside Adam ("reference per rank + gradient averaging",
SURVEY.md section 8e).  The per-rank arithmetic here is the numpy oracle; no CUDA kernel is called."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    from oracle import patchgan_oracle as orc
    from patchgan_b200 import dp
    r, w, local = dp.init_from_env('gloo')
    assert (r, w, local) == (rank, world, rank) and dp.world_size() == world and dp.rank() == rank
    assert dp.shard_seed(1234) == 1234 + rank

    # broadcast_parameters: every rank ends with rank 0's weights
    torch.manual_seed(100 + rank)
    lin = torch.nn.Linear(4, 3)
    dp.broadcast_parameters(lin)
    ref = [torch.zeros_like(p) for p in lin.parameters()]
    for t, p in zip(ref, lin.parameters()):
        t.copy_(p.data)
        dist.broadcast(t, src=0)
        assert torch.equal(t, p.data)

    # max_over_ranks (timing reduction)
    assert dp.max_over_ranks(float(rank + 1), torch.device('cpu')) == float(world)

    # one DP generator/discriminator step: per-rank oracle gradients, flat all-reduce, Adam with grad_scale = 1/world
    gk = dict(input_nc=3, output_nc=1, nf=8, activation='tanh', final_act='sigmoid')
    dk = dict(input_nc=4, ndf=8, n_layers=2, norm=False)
    G, D = orc.UNet(**gk, seed=11), orc.Discriminator(**dk, seed=12)
    tr = orc.Trainer(G, D)
    x, y = orc.synthetic_batch(1, 1, 256, seed=dp.shard_seed(1234))
    w0 = {k: v.copy() for k, v in {**G.params, **D.params}.items()}
    tr.batch(x, y, train=True)                       # fills tr.last[...] (its own Adam update is discarded below)
    names_g, names_d = list(G.params), list(D.params)
    flat_g = torch.from_numpy(np.concatenate([tr.last['gen_grads'][k].ravel() for k in names_g]))
    flat_d = torch.from_numpy(np.concatenate([tr.last['disc_grads'][k].ravel() for k in names_d]))
    hg = dp.all_reduce_sum_async(flat_g)
    hd = dp.all_reduce_sum_async(flat_d)
    hg.wait()
    hd.wait()
    np.save(os.path.join(out_dir, f'flat_g_{rank}.npy'), flat_g.numpy())
    np.save(os.path.join(out_dir, f'own_g_{rank}.npy'),
            np.concatenate([tr.last['gen_grads'][k].ravel() for k in names_g]))
    # Adam on the averaged gradients from the common starting weights
    params = {k: w0[k].copy() for k in names_g}
    opt = orc.Adam(params, 1e-3)
    off, grads = 0, {}
    for k in names_g:
        n = params[k].size
        grads[k] = (flat_g[off:off + n].numpy() / world).reshape(params[k].shape)
        off += n
    opt.step(grads)
    np.save(os.path.join(out_dir, f'w_{rank}.npy'), np.concatenate([params[k].ravel() for k in names_g]))
    dist.destroy_process_group()


def test_data_parallel_two_ranks_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    flats = [np.load(tmp_path / f'flat_g_{r}.npy') for r in range(world)]
    owns = [np.load(tmp_path / f'own_g_{r}.npy') for r in range(world)]
    ws = [np.load(tmp_path / f'w_{r}.npy') for r in range(world)]
    assert np.array_equal(flats[0], flats[1])                       # both ranks hold the same reduced buffer
    np.testing.assert_allclose(flats[0], owns[0] + owns[1], rtol=1e-6, atol=1e-9)   # == sum of the per-shard gradients
    assert not np.allclose(owns[0], owns[1])                        # the shards really differ (seed 1234 + rank)
    assert np.array_equal(ws[0], ws[1])                             # replicas stay in lock-step after the update


def test_single_process_helpers():
    from patchgan_b200 import dp
    assert dp.world_size() == 1 and dp.rank() == 0
    t = torch.ones(4)
    assert dp.all_reduce_sum_async(t).wait() and torch.equal(t, torch.ones(4))
    assert dp.max_over_ranks(3.5, torch.device('cpu')) == 3.5
    os.environ.pop('WORLD_SIZE', None)
    assert dp.init_from_env() == (0, 1, 0)
