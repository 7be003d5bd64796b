"""GPU parity of BatchNormalization / LayerNormalization (+ fused activation; dl4ds_b200/csrc/norm.cu through the
C ABI) against the oracle: the op alone (forward, input and parameter gradients, moving statistics, inference
mode), the blocks and whole networks built with ``normalization='bn' | 'ln'`` (blocks.py:63-71), and a training
step.  Tolerances as in test_gpu_engine.py (fp32 mode: forward 2e-5, gradients 2e-4 of the tensor's max; whole
networks 3e-3 on parameter gradients because of ReLU mask flips)."""
import numpy as np
import pytest
import torch

from dl4ds_b200 import SupervisedTrainer, blocks as B, nets
from dl4ds_b200.engine import Arena, Ctx
from oracle import torch_ref as R
from tests.util import compare, rel_err, trace_spec

pytestmark = pytest.mark.gpu


def _o(fn):
    def w(p, xs):
        return R._nhwc(fn(p, [R._nchw(x) for x in xs]))
    return w


@pytest.mark.parametrize('kind', ['bn', 'ln'])
@pytest.mark.parametrize('shape,act', [((3, 7, 5, 1), None), ((2, 16, 16, 8), 'relu'), ((2, 9, 11, 24), 'tanh'),
                                       ((1, 32, 32, 48), 'relu'), ((2, 6, 6, 100), 'sigmoid'),
                                       ((1, 4, 4, 256), None)])
def test_norm_op(cuda, kind, shape, act):
    fn = lambda c, xs: c.norm(xs[0], 'n', kind, act=act)
    ofn = _o(lambda p, xs: R.act(R.normalize(p, 'n', xs[0], kind), act))
    compare(fn, ofn, [shape], cuda, scale_inputs=2.0)


def test_norm_on_channel_slices_and_shared_gradient(cuda):
    """Input is a channel slice of a concat buffer and feeds two consumers (gradient accumulation path)."""
    def fn(c, xs):
        cat = c.concat([xs[0], xs[1]])
        a = c.norm(cat, 'n1', 'bn', act='relu')
        b = c.norm(cat, 'n2', 'ln', act='tanh')
        return c.add(a, b)

    def ofn(p, xs):
        cat = torch.cat(xs, dim=1)
        return torch.relu(R.normalize(p, 'n1', cat, 'bn')) + torch.tanh(R.normalize(p, 'n2', cat, 'ln'))
    compare(fn, _o(ofn), [(2, 8, 8, 8), (2, 8, 8, 4)], cuda)


def test_batchnorm_moving_statistics_and_inference(cuda):
    rng = np.random.default_rng(3)
    x = (1.5 + 2 * rng.standard_normal((4, 12, 10, 6))).astype(np.float32)
    fn = lambda c, xs: c.norm(xs[0], 'n', 'bn', act='relu')
    spec = trace_spec(fn, [x.shape])
    w = R.init_weights(spec, seed=1, bias_scale=0.1)
    arena = Arena(spec, cuda)
    arena.load({k: v.numpy() for k, v in w.items()})
    xt = torch.as_tensor(x).cuda()
    ow = {k: v.clone() for k, v in w.items()}
    for _ in range(2):                                           # two training-mode forwards
        ctx = Ctx(arena, 'fp32', training=True)
        y = fn(ctx, [ctx.input(xt)]).t.cpu().numpy()
        yo = R._nhwc(torch.relu(R.normalize(R.Params(ow), 'n', R._nchw(torch.as_tensor(x)), 'bn'))).numpy()
        assert rel_err(y, yo) <= 2e-5
    sd = arena.state_dict()
    for k in ('n/moving_mean', 'n/moving_variance'):
        assert rel_err(sd[k], ow[k].numpy()) <= 1e-5, k
        assert not np.allclose(sd[k], w[k].numpy())              # they did move
    ctx = Ctx(arena, 'fp32', training=False)                     # inference: the moving statistics
    y = fn(ctx, [ctx.input(xt)]).t.cpu().numpy()
    yo = R._nhwc(torch.relu(R.normalize(R.Params(ow, training=False), 'n', R._nchw(torch.as_tensor(x)), 'bn'))).numpy()
    assert rel_err(y, yo) <= 2e-5
    assert np.array_equal(arena.state_dict()['n/moving_mean'], sd['n/moving_mean'])


@pytest.mark.parametrize('nz', ['bn', 'ln'])
def test_blocks_with_normalization(cuda, nz):
    compare(lambda c, xs: B.conv_block(c, 'b', xs[0], 16, 'relu', True, normalization=nz),
            _o(lambda p, xs: R.conv_block(p, 'b', xs[0], 16, 'relu', True, normalization=nz)), [(2, 12, 12, 8)], cuda)
    compare(lambda c, xs: B.residual_block(c, 'b', xs[0], 16, 'relu', True, True, normalization=nz),
            _o(lambda p, xs: R.residual_block(p, 'b', xs[0], 16, 'relu', True, True, normalization=nz)),
            [(2, 12, 12, 8)], cuda)
    compare(lambda c, xs: B.dense_block(c, 'b', xs[0], 8, 'relu', False, normalization=nz),
            _o(lambda p, xs: R.dense_block(p, 'b', xs[0], 8, 'relu', False, normalization=nz)), [(2, 12, 12, 8)], cuda,
            skip_grads=('b/conv1/bias',) if nz == 'bn' else ())


def _net_case(cuda, model, ofn, batch, math='fp32'):
    shapes = [(batch,) + tuple(s) for s in model.input_shapes]
    compare(model.fn, ofn, shapes, cuda, tol=5e-5, gtol=3e-3, input_grads=False, math=math)


@pytest.mark.parametrize('nz', ['bn', 'ln'])
def test_net_resnet_spc_normalized(cuda, nz):
    m = nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (16, 16), n_blocks=3, normalization=nz)
    _net_case(cuda, m, lambda p, xs: R.net_postupsampling(p, xs, 'resnet', 'spc', 4, n_blocks=3, normalization=nz), 4)


def test_net_resnet_spc_bn_tensor_cores(cuda):
    m = nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (16, 16), n_blocks=3, normalization='bn')
    _net_case(cuda, m, lambda p, xs: R.net_postupsampling(p, xs, 'resnet', 'spc', 4, n_blocks=3, normalization='bn'),
              4, math='tf32x3')


def test_net_densenet_ln_and_unet_bn(cuda):
    m = nets.net_postupsampling('densenet', 'dc', 8, 3, 1, (8, 8), n_blocks=2, normalization='ln', attention=True)
    _net_case(cuda, m, lambda p, xs: R.net_postupsampling(p, xs, 'densenet', 'dc', 8, n_blocks=2, normalization='ln',
                                                          attention=True), 3)
    m = nets.unet_pin('unet', 2, 1, (32, 32), 1, 8, 3, normalization='bn')
    _net_case(cuda, m, lambda p, xs: R.unet_pin(p, xs, 8, 3, normalization='bn'), 2)


def test_supervised_steps_with_batchnorm(cuda):
    """Three optimizer steps of a BN network through SupervisedTrainer == the oracle's supervised_step (Adam leaves
    the moving statistics alone, the forward updates them), then predict() in inference mode."""
    np.random.seed(0)
    hr = np.random.default_rng(3).standard_normal((16, 32, 32, 1)).astype(np.float32)
    tr = SupervisedTrainer('resnet', 'spc', hr, hr[:8], hr[:8], scale=4, batch_size=8, epochs=1,
                           learning_rate=(1e-3, 1e-4), lr_decay_after=100, verbose=False, math='fp32', seed=7,
                           n_blocks=2, normalization='bn')
    tr.setup_datagen()
    tr.setup_model()
    w = {k: torch.from_numpy(v.copy()) for k, v in tr.model.get_weights().items()}
    opt = R.TFAdam(list(w), lr=R.piecewise_constant(100, 1e-3, 1e-4))
    fwd = lambda p, xs: R.net_postupsampling(p, xs, 'resnet', 'spc', 4, n_blocks=2, normalization='bn')
    for i in range(3):
        (lr,), (y,) = tr.ds_train[i % len(tr.ds_train)]
        got = tr.train_on_batch([lr], y)
        ref, _ = R.supervised_step(fwd, w, opt, [torch.from_numpy(lr)], torch.from_numpy(y))
        assert abs(got - ref) <= 2e-4 * max(1.0, abs(ref)), (i, got, ref)
    new = tr.model.get_weights()
    for k in new:
        if k.endswith(('/moving_mean', '/moving_variance')):
            assert rel_err(new[k], w[k].numpy()) <= 1e-4, k
    from tests.util import assert_adam_weights_close
    assert_adam_weights_close(new, {k: v.numpy() for k, v in w.items()}, lr=1e-3, steps=3, tight=5e-5, frac=5e-3)
    (lr,), _ = tr.ds_train[0]
    y = tr.model.predict([lr], batch_size=8)
    yo = fwd(R.Params(w, training=False), [torch.from_numpy(lr)]).detach().numpy()
    assert rel_err(y, yo) <= 1e-3
