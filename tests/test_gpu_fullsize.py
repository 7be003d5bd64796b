"""Full-size (BASELINE.json configs[1]: batch 64, 32 -> 128, residual backbone, 4x SPC) property tests: the oracle
takes minutes at this size, so the CUDA path is checked through size-independent properties instead
(linearity of the gradient in the batch, homogeneity of a bias-free linear layer, determinism of replays,
loss trajectory of the captured step against the eager step)."""
import numpy as np
import pytest
import torch

from dl4ds_b200 import SupervisedTrainer, nets
from dl4ds_b200.engine import Arena, Ctx, Var
from dl4ds_b200.spec import SpecCtx

pytestmark = pytest.mark.gpu
B, LR, S = 64, 32, 4


def _hr(n, seed=0):
    return np.random.default_rng(seed).standard_normal((n, LR * S, LR * S, 1), dtype=np.float32)


def _coarsen(hr):
    n = hr.shape[0]
    return hr.reshape(n, LR, S, LR, S, 1).mean(axis=(2, 4)).astype(np.float32)


def _grads(model, lr, hr, math):
    dev = model.arena.device
    model.arena.zero_grad()
    ctx, out = model.forward([torch.as_tensor(lr).to(dev)], training=True, math=math)
    # sum-reduced MSE-like seed: d/dy of 0.5*sum((y-t)^2) = y - t, linear in the batch
    out.grad = Var((out.buf - torch.as_tensor(hr).to(dev)).contiguous())
    ctx.backward()
    torch.cuda.synchronize()
    return model.arena.grad.clone()


@pytest.mark.parametrize('math', ['tf32x3'])
def test_gradient_is_additive_over_the_batch(cuda, math):
    """sum-loss gradients of two half batches add up to the full batch's (every wgrad kernel splits and
    atomically merges the pixel dimension differently for 32 and 64 images)."""
    m = nets.net_postupsampling('resnet', 'spc', S, 1, 0, (LR, LR), math=math).to(cuda).init_weights(seed=1)
    hr = _hr(B, 3)
    lr = _coarsen(hr)
    g_full = _grads(m, lr, hr, math)
    g_a = _grads(m, lr[:B // 2], hr[:B // 2], math)
    g_b = _grads(m, lr[B // 2:], hr[B // 2:], math)
    err = float((g_full - (g_a + g_b)).abs().max() / g_full.abs().max())
    assert err <= 2e-5, err


@pytest.mark.parametrize('cin,cout,hw', [(48, 192, 64), (48, 48, 32)])
def test_conv_homogeneity_and_superposition(cuda, cin, cout, hw):
    """conv(a*x1 + x2) == a*conv(x1) + conv(x2) for the bias-free tensor-core convolution at batch 64 (3xTF32)."""
    fn = lambda c, xs: c.conv(xs[0], 'cv', cout, k=3, bias=False)
    sc = SpecCtx()
    fn(sc, [sc.input((B, hw, hw, cin))])
    arena = Arena(sc.spec, cuda)
    arena.theta.normal_(0, 0.05)
    g = torch.Generator(device='cuda').manual_seed(0)
    x1 = torch.randn((B, hw, hw, cin), device=cuda, generator=g)
    x2 = torch.randn((B, hw, hw, cin), device=cuda, generator=g)

    def run(x):
        ctx = Ctx(arena, 'tf32x3', training=False)
        return fn(ctx, [ctx.input(x)]).buf
    y = run(2.5 * x1 + x2)
    ref = 2.5 * run(x1) + run(x2)
    assert float((y - ref).abs().max() / ref.abs().max()) <= 2e-5


def test_captured_step_replays_match_eager_and_learn(cuda):
    """20 optimizer steps at the headline size: the CUDA-graph step and the eager step produce the same loss
    trajectory (same kernels, different launch path), losses stay finite and decrease."""
    hr = _hr(2 * B, 5)
    losses = []
    for use_graph in (True, False):
        tr = SupervisedTrainer('resnet', 'spc', hr, hr[:B], hr[:B], scale=S, batch_size=B, epochs=1,
                               learning_rate=1e-3, verbose=False, seed=11)
        tr.setup_model()
        if not use_graph:
            tr.train_step.graph_fb = tr.train_step.graph_opt = None
        lr = _coarsen(hr)
        ls = [tr.train_on_batch([lr[(i % 2) * B:(i % 2 + 1) * B]], hr[(i % 2) * B:(i % 2 + 1) * B]) for i in range(20)]
        losses.append(ls)
    a, b = np.array(losses[0]), np.array(losses[1])
    assert np.all(np.isfinite(a)) and a[-1] < a[0]
    assert np.allclose(a, b, rtol=5e-4), (a, b)
