"""GPU tests of the dropout variants (blocks.py:659-706; dl4ds_b200/csrc/dropout.cu through the C ABI).
TensorFlow's random streams cannot be reproduced, so parity is checked with the masks themselves: the oracle
multiplies by the masks the CUDA generator produced (applying ``dl4ds_dropout`` to ones with the same seed, step
and layer id), and the generator is checked statistically and for determinism."""
import numpy as np
import pytest
import torch

from dl4ds_b200 import SupervisedTrainer, _lib, blocks as B, nets
from dl4ds_b200.engine import DROPOUT_KIND, Arena, Ctx
from oracle import torch_ref as R
from tests.util import compare, rel_err, trace_spec

pytestmark = pytest.mark.gpu


def gpu_mask(shape_nhwc, rate, variant, seed, step, layer_id, n_samples=None):
    n, h, w, c = shape_nhwc
    ones = torch.ones(shape_nhwc, dtype=torch.float32, device='cuda')
    out = torch.empty_like(ones)
    state = torch.tensor([seed, step], dtype=torch.int64, device='cuda')
    _lib.call('dl4ds_dropout', ones.data_ptr(), c, out.data_ptr(), c, n * h * w, h * w, n_samples or n, c, float(rate),
              DROPOUT_KIND[variant], state.data_ptr(), int(layer_id), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out.cpu()


def supplier(seed=Arena.RNG_SEED, step=1):
    """Oracle-side mask source: the i-th dropout application of the graph <-> layer id i."""
    def f(i, shape, rate, variant):
        if len(shape) == 5:       # oracle (B,T,C,H,W) <-> engine time-major frames (T*B,H,W,C), one sample = one b
            b, t, c, h, w = shape
            m = gpu_mask((t * b, h, w, c), rate, variant, seed, step, i, n_samples=b)
            return m.reshape(t, b, h, w, c).permute(1, 0, 4, 2, 3)
        n, c, h, w = shape
        return gpu_mask((n, h, w, c), rate, variant, seed, step, i).permute(0, 3, 1, 2)
    return f


def _o(fn, **kw):
    def w(p, xs):
        p.dropout_mask = supplier(**kw)
        return R._nhwc(fn(p, [R._nchw(x) for x in xs]))
    return w


def test_mask_statistics_and_determinism(cuda):
    shape, rate = (8, 64, 64, 16), 0.3
    m = gpu_mask(shape, rate, None, 11, 1, 1).numpy()
    assert set(np.unique(m)).issubset({0.0, np.float32(1 / 0.7)})
    keep = (m > 0).mean()
    assert abs(keep - 0.7) < 4 * np.sqrt(0.21 / m.size)
    assert abs(m.mean() - 1.0) < 0.01                                   # inverted dropout keeps the expectation
    g = gpu_mask(shape, rate, 'gaussian', 11, 1, 1).numpy()
    assert abs(g.mean() - 1.0) < 0.005 and abs(g.std() - np.sqrt(0.3 / 0.7)) < 0.005
    s = gpu_mask(shape, rate, 'spatial', 11, 1, 1).numpy()
    assert np.array_equal(s, np.broadcast_to(s[:, :1, :1, :], s.shape))  # one draw per (sample, channel)
    assert 0.4 < (s[:, 0, 0, :] > 0).mean() < 0.95
    # pure function of (seed, step, layer id): reproducible, and different when any of them changes
    assert np.array_equal(m, gpu_mask(shape, rate, 'vanilla', 11, 1, 1).numpy())
    for other in ((12, 1, 1), (11, 2, 1), (11, 1, 2)):
        assert (gpu_mask(shape, rate, None, *other).numpy() != m).mean() > 0.3
    # neighbouring elements are uncorrelated
    a = (m > 0).astype(np.float64).ravel()
    assert abs(np.corrcoef(a[:-1], a[1:])[0, 1]) < 0.01
    with pytest.raises(_lib.Dl4dsError):
        gpu_mask(shape, 1.0, None, 1, 1, 1)


@pytest.mark.parametrize('variant', [None, 'gaussian', 'spatial', 'mcdrop'])
def test_dropout_op_and_blocks(cuda, variant):
    fn = lambda c, xs: c.dropout(xs[0], 0.25, variant)
    compare(fn, _o(lambda p, xs: R.dropout(p, xs[0], 0.25, variant)), [(3, 9, 7, 12)], cuda)
    compare(lambda c, xs: B.conv_block(c, 'b', xs[0], 16, 'relu', True, dropout_rate=0.2, dropout_variant=variant),
            _o(lambda p, xs: R.conv_block(p, 'b', xs[0], 16, 'relu', True, dropout_rate=0.2, dropout_variant=variant)),
            [(2, 12, 12, 8)], cuda)
    compare(lambda c, xs: B.residual_block(c, 'b', xs[0], 16, 'relu', False, True, normalization='ln',
                                           dropout_rate=0.2, dropout_variant=variant),
            _o(lambda p, xs: R.residual_block(p, 'b', xs[0], 16, 'relu', False, True, normalization='ln',
                                              dropout_rate=0.2, dropout_variant=variant)), [(2, 12, 12, 8)], cuda)
    compare(lambda c, xs: B.dense_block(c, 'b', xs[0], 8, 'relu', False, dropout_rate=0.2, dropout_variant=variant),
            _o(lambda p, xs: R.dense_block(p, 'b', xs[0], 8, 'relu', False, dropout_rate=0.2,
                                           dropout_variant=variant)), [(2, 12, 12, 8)], cuda)


def test_nets_with_dropout(cuda):
    kw = dict(n_blocks=2, dropout_rate=0.2, dropout_variant='spatial')
    m = nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (16, 16), **kw)
    ofn = lambda p, xs: R.net_postupsampling(p, xs, 'resnet', 'spc', 4, **kw)

    def ofn_masked(p, xs):
        p.dropout_mask = supplier()
        return ofn(p, xs)
    compare(m.fn, ofn_masked, [(3, 16, 16, 1)], cuda, tol=5e-5, gtol=3e-3, input_grads=False)
    # the number of dropout applications (= layer ids) agrees between the builder and the oracle
    from dl4ds_b200.spec import SpecCtx
    sc = SpecCtx()
    m.fn(sc, [sc.input((1, 16, 16, 1))])
    p = R.Params()
    ofn(p, [torch.zeros(1, 16, 16, 1)])
    assert sc.n_dropout == p.n_dropout == 2 * 2 + 1 + 2

    m = nets.unet_pin('unet', 1, 1, (32, 32), 1, 8, 2, dropout_rate=0.3, dropout_variant='gaussian')

    def ofn_unet(p, xs):
        p.dropout_mask = supplier()
        return R.unet_pin(p, xs, 8, 2, dropout_rate=0.3, dropout_variant='gaussian')
    compare(m.fn, ofn_unet, [(2, 32, 32, 1), (2, 32, 32, 1)], cuda, tol=5e-5, gtol=3e-3, input_grads=False)


def test_inference_mode(cuda):
    """Plain dropout is the identity outside training; the Monte-Carlo variants stay on (blocks.py:662-677)."""
    fn0 = lambda c, xs: B.conv_block(c, 'b', xs[0], 8, 'relu')
    spec = trace_spec(fn0, [(2, 8, 8, 4)])
    w = R.init_weights(spec, seed=2, bias_scale=0.1)
    arena = Arena(spec, cuda)
    arena.load({k: v.numpy() for k, v in w.items()})
    x = torch.randn(2, 8, 8, 4, device=cuda)

    def run(variant, training):
        ctx = Ctx(arena, 'fp32', training=training)
        return B.conv_block(ctx, 'b', ctx.input(x), 8, 'relu', dropout_rate=0.5, dropout_variant=variant).t.cpu().numpy()
    base = fn0(Ctx(arena, 'fp32', training=False), [Ctx(arena, 'fp32', training=False).input(x)]).t.cpu().numpy()
    assert np.array_equal(run(None, False), base)
    assert np.array_equal(run('spatial', False), base)
    a, b = run('mcdrop', False), run('mcdrop', False)
    assert not np.allclose(a, base) and not np.allclose(a, b)        # active, and a new mask per forward pass


def test_supervised_steps_with_dropout_in_the_captured_graph(cuda):
    """Three optimizer steps (captured CUDA graph): every replay draws new masks (the RNG step is bumped by a kernel
    inside the graph), and each step equals the oracle's step under the masks of that replay."""
    np.random.seed(0)
    hr = np.random.default_rng(3).standard_normal((16, 32, 32, 1)).astype(np.float32)
    kw = dict(n_blocks=2, dropout_rate=0.2, dropout_variant='vanilla')
    tr = SupervisedTrainer('resnet', 'spc', hr, hr[:8], hr[:8], scale=4, batch_size=8, epochs=1,
                           learning_rate=(1e-3, 1e-4), lr_decay_after=100, verbose=False, math='fp32', seed=7, **kw)
    tr.setup_datagen()
    tr.setup_model()
    w = {k: torch.from_numpy(v.copy()) for k, v in tr.model.get_weights().items()}
    opt = R.TFAdam(list(w), lr=R.piecewise_constant(100, 1e-3, 1e-4))
    steps_seen = []
    for i in range(3):
        (lr,), (y,) = tr.ds_train[i % len(tr.ds_train)]
        got = tr.train_on_batch([lr], y)
        seed, step = [int(v) for v in tr.model.arena.rng_state().cpu().numpy()]
        steps_seen.append(step)

        def fwd(p, xs):
            p.dropout_mask = supplier(seed, step)
            return R.net_postupsampling(p, xs, 'resnet', 'spc', 4, **kw)
        ref, _ = R.supervised_step(fwd, w, opt, [torch.from_numpy(lr)], torch.from_numpy(y))
        assert abs(got - ref) <= 2e-4 * max(1.0, abs(ref)), (i, got, ref)
    assert steps_seen[1] == steps_seen[0] + 1 and steps_seen[2] == steps_seen[1] + 1
    from tests.util import assert_adam_weights_close
    assert_adam_weights_close(tr.model.get_weights(), {k: v.numpy() for k, v in w.items()}, lr=1e-3, steps=3,
                              tight=5e-5, frac=5e-3)


@pytest.mark.parametrize('nz,variant', [('ln', 'spatial'), ('bn', None), (None, 'gaussian')])
def test_recurrent_net_with_normalization_and_dropout(cuda, nz, variant):
    """recnet_postupsampling with RecurrentConvBlock(normalization, dropout) and the 5-D tail; the spatial variant
    is SpatialDropout3D: one draw per sample and channel over (T,H,W) of the time-major frame stack."""
    T, Bz = 3, 2
    kw = dict(n_blocks=1, normalization=nz, dropout_rate=0.2, dropout_variant=variant)
    m = nets.recnet_postupsampling('resnet', 'rc', 4, 1, 1, (8, 8), T, **kw)

    def ofn(p, xs):
        p.dropout_mask = supplier()
        x = xs[0]
        x5 = x.reshape(T, Bz, *x.shape[1:]).permute(1, 0, 2, 3, 4)      # (B,T,h,w,C)
        y5 = R.recnet_postupsampling(p, [x5, xs[1]], 'resnet', 'rc', 4, T, **kw)
        return y5.permute(1, 0, 2, 3, 4).reshape(T * Bz, *y5.shape[2:])
    compare(m.fn, ofn, [(T * Bz, 8, 8, 1), (Bz, 32, 32, 1)], cuda, tol=5e-5, gtol=3e-3, input_grads=False)


def test_spatial_dropout_over_time_major_frames(cuda):
    T, Bz = 4, 3
    m = gpu_mask((T * Bz, 6, 5, 8), 0.4, 'spatial', 5, 1, 1, n_samples=Bz).numpy().reshape(T, Bz, 6, 5, 8)
    assert np.array_equal(m, np.broadcast_to(m[:1, :, :1, :1, :], m.shape))     # shared over T, H, W
    assert len(np.unique(m[0, :, 0, 0, :])) == 2                                # but not over samples / channels
