"""One-GPU timing of the other BASELINE.json configurations at their full per-GPU sizes (configs[2..4]):
parity for these graphs is covered by tests/test_gpu_engine.py / test_gpu_api.py at reduced sizes; this script
checks that the full-size shapes run (no workspace / smem / index overflow) and records ms per optimizer step.
usage: python scratch/other_configs.py [math]"""
import json
import sys
import time
import numpy as np
import torch
sys.path.insert(0, '.')
from dl4ds_b200 import CGANTrainer, SupervisedTrainer

math = sys.argv[1] if len(sys.argv) > 1 else 'tf32x3'
rng = np.random.default_rng(1234)
out = {}


def timed(fn, n=6, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


# cfg3: densenet + attention + LCB, 8x deconv, 3 predictors + 1 static, per-GPU batch 16 (128 over 8 GPUs), LR 16 -> HR 128
B = 16
hr = rng.standard_normal((2 * B, 128, 128, 1), dtype=np.float32)
preds = [rng.standard_normal((2 * B, 128, 128, 1), dtype=np.float32) for _ in range(3)]
static = rng.standard_normal((128, 128)).astype(np.float32)
tr = SupervisedTrainer('densenet', 'dc', hr, hr[:B], hr[:B], predictors_train=preds, predictors_val=[p[:B] for p in preds],
                       predictors_test=[p[:B] for p in preds], static_vars=[static], scale=8, batch_size=B, epochs=1,
                       learning_rate=1e-3, verbose=False, math=math, seed=1, attention=True, localcon_layer=True)
tr.setup_datagen(); tr.setup_model()
x, y = tr.ds_train[0]
ms = timed(lambda: tr.train_on_batch(x, y[0]))
out['cfg3 densenet+att+LCB dc x8, 16->128, 5 LR ch + 1 aux, batch 16/GPU'] = dict(ms_per_step=ms, hr_px_per_s=B * 128 * 128 / ms * 1e3, params=tr.model.count_params())
print(out, flush=True)
del tr
torch.cuda.empty_cache()

# cfg4: recurrent resnet, 4x resize-conv, T=6, 32 -> 128, per-GPU batch 8 (32 over 4 GPUs)
B, T = 8, 6
hr = rng.standard_normal((2 * B + T, 128, 128, 1), dtype=np.float32)
tr = SupervisedTrainer('resnet', 'rc', hr, hr[:B + T], hr[:B + T], scale=4, time_window=T, batch_size=B, epochs=1,
                       learning_rate=1e-3, verbose=False, math=math, seed=1)
tr.setup_datagen(); tr.setup_model()
x, y = tr.ds_train[0]
x = [np.asarray(x[0])[..., None] if np.asarray(x[0]).ndim == 4 else x[0]] + list(x[1:])
ms = timed(lambda: tr.train_on_batch(x, y[0]))
out['cfg4 recurrent resnet rc x4, T=6, 32->128, batch 8/GPU'] = dict(ms_per_step=ms, hr_px_per_s=B * T * 128 * 128 / ms * 1e3, params=tr.model.count_params())
print(out, flush=True)
del tr
torch.cuda.empty_cache()

# cfg5: CGANTrainer, unet / pin, 256 x 256, per-GPU batch 4 (32 over 8 GPUs)
B = 4
hr = rng.standard_normal((2 * B, 256, 256, 1), dtype=np.float32)
static = rng.standard_normal((256, 256)).astype(np.float32)
from dl4ds_b200.training import cgan
from dl4ds_b200 import nets
G = nets.unet_pin('unet', 2, 1, (256, 256), 1, 8, 6, math=math).to('cuda').init_weights(seed=1)
D = nets.residual_discriminator(2, 'pin', False, 4, (256, 256), n_filters=8, n_res_blocks=4, math=math).to('cuda').init_weights(seed=2)
lr = np.concatenate([hr[:B], np.broadcast_to(static[None, :, :, None], (B, 256, 256, 1))], axis=-1).astype(np.float32)
st = np.broadcast_to(static[None, :, :, None], (B, 256, 256, 1)).astype(np.float32).copy()
go, do = cgan.Adam(2e-4, beta_1=0.5), cgan.Adam(2e-4, beta_1=0.5)
ms = timed(lambda: cgan.train_step(lr, hr[:B], G, D, go, do, gen_pxloss_function='mae', static_array=st), n=4, warm=2)
out['cfg5 cGAN unet pin 256x256, batch 4/GPU (eager step)'] = dict(ms_per_step=ms, hr_px_per_s=B * 256 * 256 / ms * 1e3,
                                                                    params=G.count_params() + D.count_params())
step = cgan.CGANStep(G, D, lr.shape, hr[:B].shape, st.shape).capture()
ms = timed(lambda: step.run(lr, hr[:B], st), n=6, warm=2)
out['cfg5 cGAN unet pin 256x256, batch 4/GPU (CUDA-graph step)'] = dict(ms_per_step=ms, hr_px_per_s=B * 256 * 256 / ms * 1e3)
print(json.dumps(out, indent=1), flush=True)
