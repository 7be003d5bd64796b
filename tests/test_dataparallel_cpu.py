"""world_size-2 gloo test (CPU) of the data-parallel exchange logic: k ranks x batch b, gradients
summed by ``step.allreduce_sum_`` and scaled by ``world_grad_scale`` (what the Adam kernel folds in),
equal the single-rank gradient of the k*b batch (MAE is a global mean => the Horovod average
identity, supervised.py:365), and the LR schedule is scaled by the world size
(supervised.py:338-352).  The gradients come from the oracle so that no GPU is needed."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dl4ds_b200 import nets
from dl4ds_b200.step import LRSchedule, allreduce_sum_, world_grad_scale
from oracle import torch_ref as R


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _grads(weights, lr, hr):
    fwd = lambda p, xs: R.net_postupsampling(p, xs, 'resnet', 'spc', 4, n_blocks=1)
    _, g = R.supervised_step(fwd, weights, None, [torch.from_numpy(lr)], torch.from_numpy(hr))
    return torch.cat([g[k].reshape(-1) for k in weights])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    m = nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (8, 8), n_blocks=1)
    w = R.init_weights(m.spec, seed=5, bias_scale=0.05)
    rng = np.random.default_rng(3)
    hr = rng.standard_normal((4, 32, 32, 1)).astype(np.float32)
    lr = hr.reshape(4, 8, 4, 8, 4, 1).mean(axis=(2, 4)).astype(np.float32)
    sl = slice(rank * 2, rank * 2 + 2)                     # batch sharding: 2 samples per rank
    flat = _grads({k: v.clone() for k, v in w.items()}, lr[sl], hr[sl])
    allreduce_sum_(flat, dist)
    flat *= world_grad_scale(dist)
    full = _grads({k: v.clone() for k, v in w.items()}, lr, hr)
    err = float((flat - full).abs().max() / full.abs().max())
    sched = LRSchedule((1e-3, 1e-4), 10, scale=float(world))
    q.put((rank, err, sched(0), sched(11)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_average_equals_full_batch():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, lr0, lr1 in res:
        assert err <= 1e-5, (rank, err)
        assert abs(lr0 - 2e-3) < 1e-12 and abs(lr1 - 2e-4) < 1e-12


def test_single_process_is_identity():
    t = torch.arange(4.0)
    assert torch.equal(allreduce_sum_(t.clone()), t) and world_grad_scale() == 1.0
