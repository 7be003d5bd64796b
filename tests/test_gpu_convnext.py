"""GPU parity of the ConvNeXt pieces (dl4ds_b200/csrc/depthwise.cu + norm.cu through the C ABI) against the oracle:
depthwise 7x7 convolution (forward, input and weight gradients), GELU, ConvNextBlock, the 'convnext' backbone in
net_postupsampling / net_pin, and activation='gelu' through the fused-epilogue ops."""
import pytest

from dl4ds_b200 import blocks as B, nets
from oracle import torch_ref as R
from tests.util import compare

pytestmark = pytest.mark.gpu


def _o(fn):
    def w(p, xs):
        return R._nhwc(fn(p, [R._nchw(x) for x in xs]))
    return w


@pytest.mark.parametrize('shape,k', [((2, 9, 11, 1), 7), ((2, 16, 16, 8), 7), ((1, 32, 32, 48), 7),
                                     ((3, 5, 4, 24), 7), ((1, 12, 12, 100), 7), ((2, 8, 8, 16), 3),
                                     ((2, 8, 8, 5), 5)])
def test_depthwise_conv(cuda, shape, k):
    c = shape[3]
    fn = lambda cx, xs: cx.depthwise_conv(xs[0], 'dw', k)
    ofn = _o(lambda p, xs: R.depthwise_conv2d(xs[0], p.get('dw/depthwise_kernel', (k, k, c, 1)),
                                              p.get('dw/bias', (c,))))
    compare(fn, ofn, [shape], cuda)


def test_gelu_and_fused_hooks(cuda):
    """activation='gelu' on a conv (+residual), an add and a normalisation: linear epilogue + standalone kernel."""
    def fn(c, xs):
        y = c.conv(xs[0], 'cv', 16, act='gelu', res=xs[1])
        y = c.norm(y, 'n', 'ln', act='gelu')
        return c.add(y, xs[1], act='gelu')

    def ofn(p, xs):
        y = R.act(R._conv(p, 'cv', xs[0], 16) + xs[1], 'gelu')
        y = R.act(R.normalize(p, 'n', y, 'ln'), 'gelu')
        return R.act(y + xs[1], 'gelu')
    compare(fn, _o(ofn), [(2, 10, 12, 8), (2, 10, 12, 16)], cuda)


@pytest.mark.parametrize('nz,proj,act', [('ln', False, 'gelu'), ('ln', True, 'relu'), ('bn', True, 'gelu')])
def test_convnext_block(cuda, nz, proj, act):
    f = 24 if proj else 16
    compare(lambda c, xs: B.convnext_block(c, 'b', xs[0], f, act, nz, use_1x1conv=proj),
            _o(lambda p, xs: R.convnext_block(p, 'b', xs[0], f, act, nz, use_1x1conv=proj)), [(2, 12, 10, 16)], cuda,
            skip_grads=_zero_bias if nz == 'bn' else ())


def _zero_bias(name):
    """A bias directly in front of a batch norm has an analytically zero gradient (noise / noise otherwise)."""
    return name.endswith('/dwconv/bias')


def _net_case(cuda, model, ofn, batch, math='fp32', skip_grads=()):
    shapes = [(batch,) + tuple(s) for s in model.input_shapes]
    compare(model.fn, ofn, shapes, cuda, tol=5e-5, gtol=3e-3, input_grads=False, math=math, skip_grads=skip_grads)


@pytest.mark.parametrize('math', ['fp32', 'tf32x3'])
def test_net_convnext_spc_with_aux(cuda, math):
    m = nets.net_postupsampling('convnext', 'spc', 4, 1, 1, (16, 16), n_blocks=3, normalization='ln',
                                activation='gelu')
    _net_case(cuda, m, lambda p, xs: R.net_postupsampling(p, xs, 'convnext', 'spc', 4, n_blocks=3, normalization='ln',
                                                          activation='gelu'), 3, math)


def test_net_convnext_pin_bn(cuda):
    m = nets.net_pin('convnext', 2, 0, (24, 24), n_blocks=2, normalization='bn')
    _net_case(cuda, m, lambda p, xs: R.net_pin(p, xs, 'convnext', n_blocks=2, normalization='bn'), 2,
              skip_grads=_zero_bias)


def test_convnext_block_layer_scale_and_drop_path(cuda):
    """ConvNextBlock(drop_path > 0, layer_scale_init_value > 0) -- blocks.py:106-129,166-169,178-183: the trainable
    per-channel gamma (forward, input gradient and d gamma) and DropPath (one draw per sample, from the Philox
    generator: the oracle multiplies by the mask the CUDA call produced)."""
    import numpy as np
    import torch
    from dl4ds_b200 import _lib
    from dl4ds_b200.engine import DROPOUT_KIND, Arena
    from tests.test_gpu_dropout import gpu_mask, supplier

    fn = lambda c, xs: B.convnext_block(c, 'cnb', xs[0], 16, 'gelu', 'ln', use_1x1conv=True, drop_path=0.3,
                                        layer_scale_init_value=1e-2)

    def ofn(p, xs):
        p.dropout_mask = supplier()
        return R._nhwc(R.convnext_block(p, 'cnb', R._nchw(xs[0]), 16, 'gelu', 'ln', use_1x1conv=True, drop_path=0.3,
                                        layer_scale_init_value=1e-2))
    compare(fn, ofn, [(6, 12, 12, 8)], cuda, tol=2e-5, gtol=3e-4)
    m = gpu_mask((64, 4, 4, 8), 0.3, 'droppath', 5, 1, 1).numpy()
    assert np.array_equal(m, np.broadcast_to(m[:, :1, :1, :1], m.shape))          # one draw per sample
    assert set(np.unique(m)).issubset({0.0, np.float32(1 / 0.7)}) and 0.5 < (m[:, 0, 0, 0] > 0).mean() < 0.9
    # Keras initialises gamma to layer_scale_init_value * ones
    from dl4ds_b200 import nets
    from dl4ds_b200.nets import Model
    mdl = Model('cnb', lambda c, xs: fn(c, xs), [(12, 12, 8)]).to(cuda).init_weights(0)
    assert np.allclose(mdl.get_weights()['cnb/gamma'], 1e-2)
