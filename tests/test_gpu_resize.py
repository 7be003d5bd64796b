"""GPU parity of the nearest / bicubic resize (`rc_interpolation`, blocks.py:457-491; dl4ds_resize_fwd / _bwd through
the C ABI) against the oracle's resampling matrices (pinned against Pillow on the CPU, tests/test_resize_cpu.py)."""
import pytest

from dl4ds_b200 import blocks as B, nets
from oracle import torch_ref as R
from tests.util import compare

pytestmark = pytest.mark.gpu


def _o(fn):
    def w(p, xs):
        return R._nhwc(fn(p, [R._nchw(x) for x in xs]))
    return w


@pytest.mark.parametrize('method', ['nearest', 'bicubic', 'bilinear', 'area', 'lanczos3', 'lanczos5', 'gaussian',
                                    'mitchellcubic'])
@pytest.mark.parametrize('shape,out', [((2, 8, 8, 3), (16, 16)), ((1, 6, 10, 8), (24, 40)), ((2, 7, 5, 1), (11, 13)),
                                       ((1, 16, 12, 4), (8, 6))])
def test_resize_op(cuda, method, shape, out):
    compare(lambda c, xs: c.resize(xs[0], out[0], out[1], method),
            _o(lambda p, xs: R.resize(xs[0], out[0], out[1], method)), [shape], cuda)


@pytest.mark.parametrize('method', ['nearest', 'bicubic', 'lanczos3', 'mitchellcubic'])
def test_resize_conv_block_and_nets(cuda, method):
    compare(lambda c, xs: B.resize_conv_block(c, 'rc', xs[0], 4, 8, method),
            _o(lambda p, xs: R.resize_conv_block(p, 'rc', xs[0], 4, 8, method)), [(2, 8, 8, 8)], cuda)
    m = nets.net_postupsampling('convnet', 'rc', 2, 2, 0, (10, 12), n_blocks=2, rc_interpolation=method)
    compare(m.fn, lambda p, xs: R.net_postupsampling(p, xs, 'convnet', 'rc', 2, n_blocks=2, rc_interpolation=method),
            [(2, 10, 12, 2)], cuda, tol=5e-5, gtol=3e-3, input_grads=False)
    m = nets.unet_pin('unet', 1, 0, (16, 16), 1, 8, 2, rc_interpolation=method)
    compare(m.fn, lambda p, xs: R.unet_pin(p, xs, 8, 2, rc_interpolation=method), [(2, 16, 16, 1)], cuda, tol=5e-5,
            gtol=3e-3, input_grads=False)
