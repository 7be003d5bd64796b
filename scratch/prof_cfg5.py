"""cfg5 (cGAN, U-Net generator 256x256 + discriminator, batch 4) eager step for an ncu launch list."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from dl4ds_b200 import nets
from dl4ds_b200.training import cgan
rng = np.random.default_rng(0)
B = 4
hr = rng.standard_normal((B, 256, 256, 1), dtype=np.float32)
static = rng.standard_normal((256, 256)).astype(np.float32)
G = nets.unet_pin('unet', 2, 1, (256, 256), 1, 8, 6, math='tf32x3').to('cuda').init_weights(seed=1)
D = nets.residual_discriminator(2, 'pin', False, 4, (256, 256), n_filters=8, n_res_blocks=4, math='tf32x3').to('cuda').init_weights(seed=2)
st = np.broadcast_to(static[None, :, :, None], (B, 256, 256, 1)).astype(np.float32).copy()
lr = np.concatenate([hr, st], axis=-1)
go, do = cgan.Adam(2e-4, beta_1=0.5), cgan.Adam(2e-4, beta_1=0.5)
for _ in range(2):
    cgan.train_step(lr, hr, G, D, go, do, gen_pxloss_function='mae', static_array=st)
torch.cuda.synchronize()
# one whole step between cudaProfilerStart/Stop (ncu --profile-from-start off)
torch.cuda.profiler.start()
cgan.train_step(lr, hr, G, D, go, do, gen_pxloss_function='mae', static_array=st)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
