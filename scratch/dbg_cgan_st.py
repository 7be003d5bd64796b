import sys; sys.path.insert(0, '.')
import numpy as np, torch
from dl4ds_b200 import nets
from dl4ds_b200.training import cgan
from oracle import torch_ref as R
cuda = 'cuda'
rng = np.random.default_rng(12)
B, T, hw = 2, 3, 8
G = nets.recnet_postupsampling('resnet', 'rc', 4, 1, 1, (hw, hw), T, n_filters=4, n_blocks=1, math='fp32').to(cuda)
D = nets.residual_discriminator(1, 'rc', True, 4, (hw, hw), n_filters=4, n_res_blocks=1, math='fp32', time_window=T).to(cuda)
gw = R.init_weights(G.spec, seed=1, bias_scale=0.05)
dw = R.init_weights(D.spec, seed=2, bias_scale=0.05)
G.set_weights({k: v.numpy() for k, v in gw.items()})
D.set_weights({k: v.numpy() for k, v in dw.items()})
lr = rng.standard_normal((B, T, hw, hw, 1)).astype(np.float32)
hr = rng.standard_normal((B, T, 4 * hw, 4 * hw, 1)).astype(np.float32)
st = rng.standard_normal((B, 4 * hw, 4 * hw, 1)).astype(np.float32)
nfeat = D.spec['dense1/kernel'][0]
masks = [(rng.random((B, 1, 1, nfeat)) < 0.6).astype(np.float32) / 0.6 for _ in range(2)]
f32 = lambda a: torch.as_tensor(a).to(cuda)
tm = lambda x: x.transpose(0, 1).reshape(x.shape[0] * x.shape[1], *x.shape[2:]).contiguous()
losses = torch.zeros(4, device=cuda)
cgan._fwd_bwd(G, D, tm(f32(lr)), tm(f32(hr)), f32(st), [f32(m) for m in masks], losses, 'mae')
torch.cuda.synchronize()
gen_fn = lambda p, xs: R.recnet_postupsampling(p, xs, 'resnet', 'rc', 4, T, n_filters=4, n_blocks=1)
disc_fn = lambda p, xs, m: R.residual_discriminator(p, xs, 'rc', 4, (hw, hw), n_filters=4, n_res_blocks=1, dropout_mask=m, is_spatiotemporal=True)
ref, gg, dg = R.cgan_step(gen_fn, disc_fn, gw, dw, None, None, torch.from_numpy(lr), torch.from_numpy(hr), torch.from_numpy(st),
                          mask_real=torch.from_numpy(masks[0].reshape(B, nfeat)), mask_fake=torch.from_numpy(masks[1].reshape(B, nfeat)))
print('losses', losses.cpu().numpy(), ref)
for name, model, gr in (('G', G, gg), ('D', D, dg)):
    mine = model.arena.grads()
    for k in gr:
        a, b = mine[k], gr[k].numpy()
        e = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
        print('%s %-50s max|ref| %.3e rel err %.3e %s' % (name, k, np.abs(b).max(), e, 'BAD' if e > 3e-3 else ''))
