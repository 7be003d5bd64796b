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
