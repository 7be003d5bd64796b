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
# ---- BASELINE config 3 inputs (3 predictors + 1 static field, densenet + attention + LCB, 8x deconvolution, batch 16
#      per GPU): one epoch through SupervisedTrainer.run() from the host generator (per-sample numpy / cv2 loop, as
#      the reference) and from the device-resident generator
B3, N3 = 16, 16 * 12
hr3 = rng.standard_normal((N3 + 2 * B3, 128, 128, 1), dtype=np.float32)
preds = [rng.standard_normal((N3 + 2 * B3, 128, 128, 1), dtype=np.float32) for _ in range(3)]
static = rng.standard_normal((128, 128)).astype(np.float32)
data = {}
for on_dev in (False, True):
    np.random.seed(0)
    tr = SupervisedTrainer('densenet', 'dc', hr3[:N3], hr3[N3:N3 + B3], hr3[N3 + B3:],
                           predictors_train=[p[:N3] for p in preds], predictors_val=[p[N3:N3 + B3] for p in preds],
                           predictors_test=[p[N3 + B3:] for p in preds], static_vars=[static.copy()], scale=8,
                           batch_size=B3, epochs=2, learning_rate=1e-3, verbose=False, math=math, seed=1,
                           attention=True, localcon_layer=True, data_on_device=on_dev)
    tr.setup_datagen()
    tr.setup_model()
    gen = tr.ds_train
    st = tr.train_step

    def epoch():
        if on_dev:
            for _ in tr.train_on_device_batches(gen, list(range(len(gen)))):
                pass
        else:
            for _ in tr.train_on_batches(((gen[i][0], gen[i][1][0]) for i in range(len(gen)))):
                pass
    epoch()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    epoch()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / len(gen) * 1e3
    data['device generator' if on_dev else 'host generator'] = dict(
        ms_per_step=round(ms, 3), hr_px_per_s=round(B3 * 128 * 128 / ms * 1e3), steps=len(gen))
    print('cfg3 data path', on_dev, data, file=sys.stderr, flush=True)
    del tr
    torch.cuda.empty_cache()
out['cfg3 (densenet+att+LCB dc x8, 3 predictors + 1 static, batch 16): epoch incl. batch creation'] = data

print(json.dumps({'math': math, 'batch': B, 'timing': 'wall clock around train_on_batch (host batch -> loss float), '
                  '10 steps after 3 warm-up', 'rows': out}, indent=1))
