"""One-GPU timing of the headline model (resnet + 4x SPC, 32 -> 128, batch 64) with the SURVEY 8f row-3 options
switched on, at full size: SSIM-family losses, batch / layer norm, dropout, and the convnext backbone.  Parity of
each is covered at reduced sizes in tests/test_gpu_{losses,norm,dropout,convnext}.py; this records ms per optimizer
step through SupervisedTrainer.train_on_batch (host batches, captured graph).
usage: python scratch/next_configs.py [math] > profiles/<round>_next_configs.json"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from dl4ds_b200 import SupervisedTrainer   # noqa: E402

math = sys.argv[1] if len(sys.argv) > 1 else 'tf32x3'
rng = np.random.default_rng(1234)
B = 64
hr = rng.standard_normal((2 * B, 128, 128, 1), dtype=np.float32)
out = {}


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


CASES = [
    ('resnet spc x4, MAE (headline)', 'resnet', {}),
    ('resnet spc x4, loss=dssim_mae', 'resnet', dict(loss='dssim_mae')),
    ('resnet spc x4, loss=msdssim_mae_mse', 'resnet', dict(loss='msdssim_mae_mse')),
    ('resnet spc x4, normalization=bn', 'resnet', dict(normalization='bn')),
    ('resnet spc x4, normalization=ln', 'resnet', dict(normalization='ln')),
    ('resnet spc x4, dropout 0.2 spatial', 'resnet', dict(dropout_rate=0.2, dropout_variant='spatial')),
    ('convnext spc x4, ln, gelu', 'convnext', dict(normalization='ln', activation='gelu')),
]
for label, backbone, kw in CASES:
    np.random.seed(0)
    tr = SupervisedTrainer(backbone, 'spc', hr, hr[:B], hr[:B], scale=4, batch_size=B, epochs=1, learning_rate=1e-3,
                           verbose=False, math=math, seed=1, **kw)
    tr.setup_datagen()
    tr.setup_model()
    x, y = tr.ds_train[0]
    ms = timed(lambda: tr.train_on_batch(x, y[0]))
    out[label] = dict(ms_per_step=round(ms, 3), hr_px_per_s=round(B * 128 * 128 / ms * 1e3), params=tr.model.count_params(),
                      launches_per_step=tr.train_step.launches_per_step)
    print(label, out[label], file=sys.stderr, flush=True)
    del tr
    torch.cuda.empty_cache()
print(json.dumps({'math': math, 'batch': B, 'timing': 'wall clock around train_on_batch (host batch -> loss float), '
                  '10 steps after 3 warm-up', 'rows': out}, indent=1))
