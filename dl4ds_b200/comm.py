"""Data-parallel communicator: the ``dl4ds_comm_*`` C ABI (NCCL over NVLink / NVSwitch) behind the exchange step
of the hot path -- the gradient all-reduce and the rank-0 broadcast that the reference delegates to Horovod
(training/supervised.py:363-369, training/cgan.py:608-637).

One communicator per process (one process per GPU).  The launcher's process group (``torch.distributed``: NCCL
on GPUs) is used ONCE, to hand rank 0's NCCL unique id to the other ranks; after that every collective of the
training step is ``dl4ds_comm_allreduce_sum`` / ``dl4ds_comm_broadcast`` on the step's own stream, so the exchange
is captured inside the step's CUDA graph (forward, backward, all-reduce, Adam = one graph, no host gap between
them).  CPU tensors (the gloo tests of the host logic) fall back to ``torch.distributed``.
"""
import ctypes

import torch

from . import _lib

_state = {'ready': False, 'size': 1, 'rank': 0}


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def ensure(dist=None):
    """Create the process-wide NCCL communicator (idempotent).  Returns True when a multi-rank communicator exists."""
    if _state['ready']:
        return _state['size'] > 1
    d = dist if dist is not None else _dist()
    if d is None or not torch.cuda.is_available():
        return False
    lib = _lib.load()
    nbytes = lib.dl4ds_comm_unique_id_bytes()
    rank, world = d.get_rank(), d.get_world_size()
    buf = (ctypes.c_ubyte * nbytes)()
    if rank == 0:
        _lib.call('dl4ds_comm_get_unique_id', ctypes.addressof(buf))
    box = [bytes(buf)]
    d.broadcast_object_list(box, src=0)            # the only use of the launcher's group on this path
    ctypes.memmove(buf, box[0], nbytes)
    _lib.call('dl4ds_comm_init_rank', ctypes.addressof(buf), world, rank)
    _state.update(ready=True, size=world, rank=rank)
    return world > 1


def size():
    return _state['size'] if _state['ready'] else 1


def is_ready():
    return _state['ready'] and _state['size'] > 1


def allreduce_sum_(flat):
    """In-place sum over the ranks of a flat fp32 CUDA tensor, on the current stream (capturable)."""
    assert flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous()
    _lib.call('dl4ds_comm_allreduce_sum', flat.data_ptr(), flat.numel(), torch.cuda.current_stream().cuda_stream)
    return flat


def broadcast_(t, root=0):
    """In-place broadcast of a contiguous CUDA tensor from ``root``, on the current stream."""
    assert t.is_cuda and t.is_contiguous()
    _lib.call('dl4ds_comm_broadcast', t.data_ptr(), t.numel() * t.element_size(), int(root),
              torch.cuda.current_stream().cuda_stream)
    return t


def destroy():
    if _state['ready']:
        torch.cuda.synchronize()
        _lib.call('dl4ds_comm_destroy')
        _state.update(ready=False, size=1, rank=0)
