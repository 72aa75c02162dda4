"""Data-parallel plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on B200s, gloo in the CPU
tests).  The reference is single-process; batches shard on the batch axis because the only
normalisation is per-sample InstanceNorm, so the only exchange is one gradient all-reduce per optimizer.

Semantics: each rank runs the reference step on its own shard and the flat gradient buffers are averaged
("reference per rank + gradient averaging", i.e. DDP semantics).
"""
import os

import torch
import torch.distributed as dist


def is_dist():
    return dist.is_available() and dist.is_initialized()


def world_size():
    return dist.get_world_size() if is_dist() else 1


def rank():
    return dist.get_rank() if is_dist() else 0


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK /
    MASTER_ADDR / MASTER_PORT).  Returns (rank, world, local_rank).  No-op for a single process."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1:
        return 0, 1, 0
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
        torch.cuda.set_device(local)
    if not is_dist():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        kw = {}
        if backend == 'nccl':
            kw['device_id'] = torch.device('cuda', local)
        dist.init_process_group(backend=backend, **kw)
    return dist.get_rank(), dist.get_world_size(), local


class _Done:
    def wait(self):
        return True


def all_reduce_sum_async(flat):
    """Sum-all-reduce one flat gradient buffer in place; returns a handle whose ``wait()`` orders the current
    stream after the reduction.  (The 1/world factor is folded into the Adam kernel's grad_scale.)"""
    if world_size() == 1:
        return _Done()
    return dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)


def broadcast_parameters(module, src=0):
    """Make every rank start from rank `src`'s weights (the reference has no notion of ranks)."""
    if world_size() == 1:
        return
    for p in module.parameters():
        dist.broadcast(p.data, src=src)


def shard_seed(base_seed):
    """Per-rank data seed: SURVEY.md section 8(d): data seed = 1234 + rank."""
    return base_seed + rank()


def mean_over_ranks(value, device):
    """Mean of a python float over ranks (epoch metrics that drive a learning-rate scheduler)."""
    if world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item()) / world_size()


def max_over_ranks(value, device):
    """Max of a python float over ranks (used for timing: the job is as slow as its slowest rank)."""
    if world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
