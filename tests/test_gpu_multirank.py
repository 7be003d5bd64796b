"""2-rank NCCL test (needs >= 2 GPUs, otherwise skipped): two ranks x batch 4 through
SupervisedTrainer.train_on_batch end with the same weights as one rank x batch 8 with the same
samples (gradient all-reduce + 1/world in Adam + LR x world, supervised.py:338-369)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from dl4ds_b200 import SupervisedTrainer
    rng = np.random.default_rng(0)
    hr = rng.standard_normal((8, 64, 64, 1)).astype(np.float32)
    lr = hr.reshape(8, 16, 4, 16, 4, 1).mean(axis=(2, 4)).astype(np.float32)
    b = 8 // world
    tr = SupervisedTrainer('resnet', 'spc', hr, hr, hr, scale=4, batch_size=b, epochs=1, learning_rate=1e-3 / world,
                           verbose=False, math='fp32', seed=1 + rank, n_blocks=2)   # different seeds: broadcast must fix it
    tr.setup_model()
    sl = slice(rank * b, (rank + 1) * b)
    for _ in range(3):
        tr.train_on_batch([lr[sl]], hr[sl])
    torch.cuda.synchronize()
    if rank == 0:
        np.savez(out_path, **{k.replace('/', '|'): v for k, v in tr.model.get_weights().items()})
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def test_two_ranks_equal_one_rank(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    outs = []
    for world in (1, 2):
        port = _free_port()
        out = str(tmp_path / ('w%d.npz' % world))
        procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=300)
            assert p.exitcode == 0
        outs.append(dict(np.load(out)))
    from tests.util import assert_adam_weights_close
    # 2 ranks x batch b vs 1 rank x batch 2b: same gradients up to summation order (see assert_adam_weights_close)
    assert_adam_weights_close(outs[0], outs[1], lr=1e-3, steps=3, tight=2e-5)
