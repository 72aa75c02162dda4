"""Data-parallel plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on B200s, gloo in the CPU
tests).  The reference is single-process; batches shard on the batch axis because the only
normalisation is per-sample InstanceNorm, so the only exchange is one gradient all-reduce per optimizer.

Semantics: each rank runs the reference step on its own shard and the flat gradient buffers are averaged
("reference per rank + gradient averaging", i.e. DDP semantics).
"""
import ctypes
import glob
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


# ------------------------------------------------------------------------------------------------
# Our own NCCL communicator (the torch-bundled libnccl, bound with ctypes).  The process group's collectives could not be
# captured into the step's CUDA graph on this stack (the capture hung), raw ncclAllReduce calls on a communicator of our
# own can: the gradient all-reduces become nodes of the step's graph, on the stream whose work they follow, and overlap
# the backward kernels of the other network.  torch.distributed still does the rendezvous (it carries the NCCL unique id).
# ------------------------------------------------------------------------------------------------
class _UniqueId(ctypes.Structure):
    _fields_ = [('internal', ctypes.c_byte * 128)]


NCCL_FLOAT32, NCCL_SUM = 7, 0
# SMs left to the NCCL kernels while they overlap the backward (= NCCL_MAX_CTAS).  Measured on B200 (cfg 3, ms/step): 16 vs 32 SMs
# at 2 GPUs 2.54 / 2.60, at 4 GPUs 2.60 / 2.63, at 8 GPUs 2.68 / 2.63 -- the 42 MB all-reduce takes 176 us at 2 GPUs and 259 us
# at 8 with 16 CTAs, so only the 8-GPU ring is worth more SMs.  PATCHGAN_B200_NCCL_SMS overrides.
NCCL_SMS = 16
_COMM = {'tried': False, 'lib': None, 'comm': None}


def _load_nccl():
    base = os.path.dirname(os.path.dirname(torch.__file__))
    cands = glob.glob(os.path.join(base, 'nvidia', 'nccl', 'lib', 'libnccl.so*')) + ['libnccl.so.2']
    for c in cands:
        try:
            lib = ctypes.CDLL(c)
        except OSError:
            continue
        lib.ncclGetUniqueId.argtypes = [ctypes.POINTER(_UniqueId)]
        lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _UniqueId, ctypes.c_int]
        lib.ncclAllReduce.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_void_p]
        lib.ncclGetErrorString.restype = ctypes.c_char_p
        lib.ncclGetErrorString.argtypes = [ctypes.c_int]
        return lib
    return None


def raw_comm():
    """The process-wide raw NCCL communicator (created on first use, collectively: every rank must get here), or None when
    the job is not an NCCL job (single process, gloo) or PATCHGAN_B200_RAW_NCCL=0."""
    if _COMM['tried']:
        return _COMM['comm']
    _COMM['tried'] = True
    if world_size() == 1 or dist.get_backend() != 'nccl' or os.environ.get('PATCHGAN_B200_RAW_NCCL', '1') == '0':
        return None
    lib = _load_nccl()
    if lib is None:
        return None
    # The collectives run beside the backward kernels: cap NCCL's CTAs, and keep as many SMs out of the grid-barrier
    # kernels' reach (see pg_set_sm_limit), so that neither can starve the other of the SMs it needs to make progress.
    global NCCL_SMS
    NCCL_SMS = int(os.environ.get('PATCHGAN_B200_NCCL_SMS', '32' if world_size() >= 8 else '16'))
    os.environ.setdefault('NCCL_MAX_CTAS', str(NCCL_SMS))
    from . import _lib as L
    nsm = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    L.check(L.lib().pg_set_sm_limit(max(nsm - NCCL_SMS, nsm // 2)), 'pg_set_sm_limit')
    _COMM['nsm'] = nsm
    uid = _UniqueId()
    if rank() == 0:
        rc = lib.ncclGetUniqueId(ctypes.byref(uid))
        if rc != 0:
            raise RuntimeError(f'ncclGetUniqueId: {lib.ncclGetErrorString(rc).decode()}')
    dev = torch.device('cuda', torch.cuda.current_device())
    t = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    ctypes.memmove(ctypes.byref(uid), bytes(t.cpu().tolist()), 128)
    comm = ctypes.c_void_p()
    rc = lib.ncclCommInitRank(ctypes.byref(comm), world_size(), uid, rank())
    if rc != 0:
        raise RuntimeError(f'ncclCommInitRank: {lib.ncclGetErrorString(rc).decode()}')
    _COMM['lib'], _COMM['comm'] = lib, comm
    return comm


def reserve_sms(on):
    """Keep NCCL_SMS SMs out of the reach of the one-launch conv + InstanceNorm kernels planned from now on (on=True: a raw
    all-reduce may be in flight beside them) or give them the whole GPU (on=False: the caller guarantees that none is).
    The limit is host-side planning state: it is baked into the launches captured after the call."""
    if _COMM.get('comm') is None:
        return
    from . import _lib as L
    nsm = _COMM['nsm']
    L.check(L.lib().pg_set_sm_limit(max(nsm - NCCL_SMS, nsm // 2) if on else 0), 'pg_set_sm_limit')


def raw_all_reduce_sum_(flat, first=0, count=None):
    """In-place fp32 sum-all-reduce of flat[first : first + count] on the CURRENT stream through the raw communicator
    (graph-capturable; no host synchronisation)."""
    comm = raw_comm()
    if comm is None:
        raise RuntimeError('raw NCCL communicator unavailable')
    if flat.dtype != torch.float32 or not flat.is_contiguous():
        raise RuntimeError('raw_all_reduce_sum_: contiguous float32 buffers only')
    n = flat.numel() - first if count is None else count
    if n <= 0:
        return
    ptr = flat.data_ptr() + first * 4
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = _COMM['lib'].ncclAllReduce(ptr, ptr, n, NCCL_FLOAT32, NCCL_SUM, comm, st)
    if rc != 0:
        raise RuntimeError(f'ncclAllReduce: {_COMM["lib"].ncclGetErrorString(rc).decode()}')


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
