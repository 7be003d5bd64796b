"""GPU parity tests (run with ``-m gpu`` on the B200 box): every engine op and every network graph,
forward AND backward, through the C ABI (libdl4ds_b200.so) against the oracle
(oracle/torch_ref.py -- torch-CPU fp32 restatement of the reference's Keras graphs; parity
unpinned by the reference, see oracle/__init__.py).

Tolerances (fp32 CUDA-core math mode, the exact-fp32 anchor): forward <= 2e-5 * max|y|,
gradients <= 2e-4 * max|g| (fp32 summation-order noise of split-K atomics over up to 1e6 pixels).
"""
import numpy as np
import pytest
import torch

from dl4ds_b200 import blocks as B
from dl4ds_b200 import nets
from dl4ds_b200.spec import SpecCtx
from oracle import torch_ref as R
from tests.util import compare

pytestmark = pytest.mark.gpu


def _o(fn):
    """Wrap an oracle NCHW block function into NHWC-in / NHWC-out."""
    def w(p, xs):
        return R._nhwc(fn(p, [R._nchw(x) for x in xs]))
    return w


# ------------------------------------------------------------------------------------------ conv
@pytest.mark.parametrize('shape,cout,k,act,stride,padding', [
    ((2, 8, 8, 1), 8, 3, None, 1, 'same'),
    ((2, 9, 7, 3), 5, 3, 'relu', 1, 'same'),
    ((1, 16, 16, 8), 16, 3, 'relu', 1, 'same'),
    ((2, 12, 12, 24), 40, 3, None, 1, 'same'),
    ((2, 10, 10, 48), 48, 3, 'tanh', 1, 'same'),
    ((2, 6, 6, 16), 64, 1, 'sigmoid', 1, 'same'),
    ((2, 11, 11, 4), 6, 5, None, 1, 'same'),
    ((2, 16, 16, 8), 8, 3, None, 2, 'same'),
    ((2, 15, 13, 8), 8, 3, 'relu', 2, 'same'),
    ((2, 17, 17, 8), 8, 3, None, 2, 'valid'),
    ((1, 20, 20, 2), 3, 7, None, 1, 'same'),
    ((3, 1, 1, 16), 32, 1, 'sigmoid', 1, 'same'),
])
def test_conv(cuda, shape, cout, k, act, stride, padding):
    fn = lambda c, xs: c.conv(xs[0], 'cv', cout, k=k, act=act, stride=stride, padding=padding)
    ofn = _o(lambda p, xs: R.act(R._conv(p, 'cv', xs[0], cout, k=k, stride=stride, padding=padding), act))
    compare(fn, ofn, [shape], cuda)


def test_conv_residual_epilogue(cuda):
    def fn(c, xs):
        return c.conv(xs[0], 'cv', 12, act='relu', res=xs[1])
    ofn = _o(lambda p, xs: R.act(R._conv(p, 'cv', xs[0], 12) + xs[1], 'relu'))
    compare(fn, ofn, [(2, 9, 10, 7), (2, 9, 10, 12)], cuda)


@pytest.mark.parametrize('r,c', [(2, 12), (2, 48), (5, 2), (3, 4)])
def test_conv_depth_to_space(cuda, r, c):
    fn = lambda cx, xs: cx.conv(xs[0], 'cv', c * r * r, d2s=r)
    ofn = _o(lambda p, xs: R.depth_to_space(R._conv(p, 'cv', xs[0], c * r * r), r))
    compare(fn, ofn, [(2, 6, 5, 8)], cuda)


@pytest.mark.parametrize('shape,cout,stride', [((2, 4, 4, 8), 6, 2), ((1, 5, 3, 4), 4, 2),
                                               ((1, 3, 3, 4), 5, 4), ((2, 8, 8, 48), 48, 2)])
def test_conv_transpose(cuda, shape, cout, stride):
    fn = lambda c, xs: c.conv_transpose(xs[0], 'ct', cout, 9, stride, act='relu')

    def ofn_(p, xs):
        w = p.get('ct/kernel', (9, 9, cout, xs[0].shape[1]))
        return R.act(R.conv2d_transpose_same(xs[0], w, stride), 'relu')
    compare(fn, _o(ofn_), [shape], cuda)


def test_shared_conv_accumulates(cuda):
    """SubpixelConvolutionBlock applies ONE conv2x at every x2 stage (blocks.py:415,421-422)."""
    fn = lambda c, xs: B.subpixel_block(c, 'spc', xs[0], 4, 8)
    ofn = _o(lambda p, xs: R.subpixel_block(p, 'spc', xs[0], 4, 8))
    compare(fn, ofn, [(2, 5, 6, 8)], cuda)


# ------------------------------------------------------------------------------------------ blocks
@pytest.mark.parametrize('att,proj', [(False, False), (False, True), (True, True)])
def test_residual_block(cuda, att, proj):
    cin = 16 if not proj else 8
    fn = lambda c, xs: B.residual_block(c, 'rb', xs[0], 16, 'relu', att, proj)
    ofn = _o(lambda p, xs: R.residual_block(p, 'rb', xs[0], 16, 'relu', att, proj))
    compare(fn, ofn, [(2, 8, 8, cin)], cuda)


@pytest.mark.parametrize('att', [False, True])
def test_dense_block(cuda, att):
    fn = lambda c, xs: B.dense_block(c, 'db', xs[0], 8, 'relu', att)
    ofn = _o(lambda p, xs: R.dense_block(p, 'db', xs[0], 8, 'relu', att))
    compare(fn, ofn, [(2, 7, 9, 12)], cuda)


def test_conv_block_attention(cuda):
    fn = lambda c, xs: B.conv_block(c, 'cb', xs[0], 8, None, True)
    ofn = _o(lambda p, xs: R.conv_block(p, 'cb', xs[0], 8, None, True))
    compare(fn, ofn, [(3, 16, 16, 8)], cuda)


def test_channel_attention_wide(cuda):
    fn = lambda c, xs: c.channel_attention(xs[0], 'att')
    ofn = _o(lambda p, xs: R.channel_attention(p, 'att', xs[0], 80))
    compare(fn, ofn, [(2, 6, 6, 80)], cuda)


def test_localized_conv_block(cuda):
    fn = lambda c, xs: B.localized_conv_block(c, 'lcb', xs[0], 2)
    ofn = _o(lambda p, xs: R.localized_conv_block(p, 'lcb', xs[0], 2))
    compare(fn, ofn, [(3, 8, 6, 10)], cuda)


def test_resize_conv_block(cuda):
    fn = lambda c, xs: B.resize_conv_block(c, 'rc', xs[0], 4, 8)
    ofn = _o(lambda p, xs: R.resize_conv_block(p, 'rc', xs[0], 4, 8))
    compare(fn, ofn, [(2, 5, 7, 8)], cuda)


def test_resize_bilinear_down(cuda):
    fn = lambda c, xs: c.conv(c.resize_bilinear(xs[0], 4, 5), 'cv', 3, k=1)
    ofn = _o(lambda p, xs: R._conv(p, 'cv', R.resize_bilinear(xs[0], 4, 5), 3, k=1))
    compare(fn, ofn, [(2, 12, 15, 3)], cuda)


def test_maxpool_padconcat(cuda):
    def fn(c, xs):
        y = c.conv(xs[0], 'cv', 6)
        d = c.maxpool2(y)
        u = c.resize_bilinear(d, d.H * 2, d.W * 2)
        return c.conv(B.pad_concat(c, u, y), 'cv2', 4)

    def ofn_(p, xs):
        y = R._conv(p, 'cv', xs[0], 6)
        d = R.maxpool2(y)
        u = R.resize_bilinear(d, d.shape[2] * 2, d.shape[3] * 2)
        return R._conv(p, 'cv2', R.pad_concat(u, y), 4)
    compare(fn, _o(ofn_), [(2, 9, 11, 3)], cuda)


def test_deconv_block_x8(cuda):
    fn = lambda c, xs: B.deconv_block(c, 'dc', xs[0], 8, 6, 'relu')
    ofn = _o(lambda p, xs: R.deconv_block(p, 'dc', xs[0], 8, 6, 'relu'))
    compare(fn, ofn, [(1, 3, 4, 4)], cuda)


def test_convlstm_block(cuda):
    T, Bz = 3, 2

    def fn(c, xs):
        return B.recurrent_conv_block(c, 'rcb', xs[0], 4, T, 'relu')

    def ofn(p, xs):     # xs[0]: time-major frames (T*B,H,W,C) NHWC
        x = xs[0]
        x5 = x.reshape(T, Bz, *x.shape[1:]).permute(1, 0, 4, 2, 3)      # (B,T,C,H,W)
        y5 = R.recurrent_conv_block(p, 'rcb', x5, 4, 'relu')            # (B,T,F,H,W)
        return y5.permute(1, 0, 3, 4, 2).reshape(T * Bz, y5.shape[3], y5.shape[4], y5.shape[2])
    compare(fn, ofn, [(T * Bz, 6, 5, 3)], cuda, gtol=5e-4)


# ------------------------------------------------------------------------------------------ nets
def _net_case(cuda, model, ofn, batch, tol=5e-5, gtol=3e-3):
    """Whole-network parity.  Forward stays at 5e-5; the parameter-gradient bound is 3e-3 of each tensor's max: a
    ReLU (block activations, the squeeze MLP of channel attention) whose pre-activation differs by ~1e-7 between the
    two summation orders flips its mask and moves a weight gradient by ~1e-3 of its max, and the split-K atomics
    make that order vary from run to run (observed 5.6e-4 on one of many runs).  The per-op tests hold 2e-4."""
    shapes = []
    for s in model.input_shapes:
        if len(s) == 4:
            shapes.append((batch * s[0],) + tuple(s[1:]))
        else:
            shapes.append((batch,) + tuple(s))
    compare(model.fn, ofn, shapes, cuda, tol=tol, gtol=gtol, input_grads=False)


def test_net_resnet_spc_cfg1(cuda):
    """BASELINE config 1/2 graph: resnet + 4x SPC, 32->128, 1 channel (batch 4)."""
    m = nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (32, 32))
    assert m.count_params() == 204405
    ofn = lambda p, xs: R.net_postupsampling(p, xs, 'resnet', 'spc', 4)
    _net_case(cuda, m, ofn, 4)


def test_net_densenet_dc_lcb_cfg3(cuda):
    """BASELINE config 3 graph: densenet + attention + LCB, 8x deconv, 5 LR ch + 1 HR aux."""
    m = nets.net_postupsampling('densenet', 'dc', 8, 5, 1, (8, 8), attention=True, localcon_layer=True)
    ofn = lambda p, xs: R.net_postupsampling(p, xs, 'densenet', 'dc', 8, attention=True, localcon_layer=True)
    _net_case(cuda, m, ofn, 3)


def test_net_convnet_rc(cuda):
    m = nets.net_postupsampling('convnet', 'rc', 2, 2, 0, (10, 12), n_blocks=2)
    ofn = lambda p, xs: R.net_postupsampling(p, xs, 'convnet', 'rc', 2, n_blocks=2)
    _net_case(cuda, m, ofn, 2)


def test_net_pin(cuda):
    m = nets.net_pin('resnet', 2, 1, (16, 16), n_blocks=3)
    ofn = lambda p, xs: R.net_pin(p, xs, 'resnet', n_blocks=3)
    _net_case(cuda, m, ofn, 2)


def test_net_unet_pin_cfg5(cuda):
    """BASELINE config 5 generator graph (reduced grid): unet / pin, 2 in ch + 1 aux."""
    m = nets.unet_pin('unet', 2, 1, (32, 32), 1, 8, 6)
    ofn = lambda p, xs: R.unet_pin(p, xs, 8, 6)
    _net_case(cuda, m, ofn, 2)


def test_net_recnet_cfg4(cuda):
    """BASELINE config 4 graph (reduced): recurrent resnet + 4x resize-conv, T=3; includes the
    5-D channel-attention (T,H)-pooling quirk (blocks.py:587)."""
    T, Bz = 3, 2
    m = nets.recnet_postupsampling('resnet', 'rc', 4, 1, 1, (8, 8), T, n_blocks=1)

    def ofn(p, xs):
        x = xs[0]
        x5 = x.reshape(T, Bz, *x.shape[1:]).permute(1, 0, 2, 3, 4)      # (B,T,h,w,C)
        y5 = R.recnet_postupsampling(p, [x5, xs[1]], 'resnet', 'rc', 4, T, n_blocks=1)
        return y5.permute(1, 0, 2, 3, 4).reshape(T * Bz, *y5.shape[2:])
    shapes = [(T * Bz, 8, 8, 1), (Bz, 32, 32, 1)]
    compare(m.fn, ofn, shapes, cuda, tol=5e-5, gtol=3e-3, input_grads=False)


@pytest.mark.parametrize('ups,scale', [('pin', 4), ('spc', 4), ('spc', 2)])
def test_discriminator(cuda, ups, scale):
    lr = (20, 20) if ups == 'pin' else (5, 5)
    m = nets.residual_discriminator(2, ups, False, scale, lr, n_res_blocks=2)
    mask = (np.random.default_rng(5).random((3, 1, 1, 16)) > 0.4).astype(np.float32) / 0.6

    def fn(c, xs):
        mk = c.input((3, 1, 1, 16)) if isinstance(c, SpecCtx) else c.input(torch.as_tensor(mask).cuda())
        return m.fn(c, [xs[0], xs[1], mk])

    def ofn(p, xs):
        y = R.residual_discriminator(p, xs, ups, scale, lr, n_res_blocks=2,
                                     dropout_mask=torch.as_tensor(mask.reshape(3, 16)))
        return y.reshape(3, 1, 1, 1)
    shapes = [(3,) + s for s in m.input_shapes]
    compare(fn, ofn, shapes, cuda, tol=5e-5, gtol=1e-3)


@pytest.mark.parametrize('ups,scale,math', [('pin', 4, 'fp32'), ('spc', 4, 'fp32'), ('pin', 4, 'tf32x3')])
def test_discriminator_spatiotemporal(cuda, ups, scale, math):
    """Spatio-temporal residual_discriminator (discriminator.py:25-33,42-47,73-74): RecurrentConvBlock('ln') stem on
    the LR branch, per-frame Conv2D / ResidualBlocks, GlobalAveragePooling3D, against the oracle's 5-D statement."""
    T, Bz = 3, 2
    lr = (16, 16) if ups == 'pin' else (8, 8)
    m = nets.residual_discriminator(2, ups, True, scale, lr, n_res_blocks=1, time_window=T)
    nfeat = m.spec['dense1/kernel'][0]
    mask = (np.random.default_rng(5).random((Bz, 1, 1, nfeat)) > 0.4).astype(np.float32) / 0.6

    def fn(c, xs):
        mk = c.input((Bz, 1, 1, nfeat)) if isinstance(c, SpecCtx) else c.input(torch.as_tensor(mask).cuda())
        return m.fn(c, [xs[0], xs[1], mk])

    def five(x):
        return x.reshape(T, Bz, *x.shape[1:]).permute(1, 0, 2, 3, 4)          # time-major frames -> (B,T,H,W,C)

    def ofn(p, xs):
        y = R.residual_discriminator(p, [five(xs[0]), five(xs[1])], ups, scale, lr, n_res_blocks=1,
                                     dropout_mask=torch.as_tensor(mask.reshape(Bz, nfeat)), is_spatiotemporal=True)
        return y.reshape(Bz, 1, 1, 1)
    shapes = [(T * Bz,) + tuple(s[1:]) for s in m.input_shapes]
    compare(fn, ofn, shapes, cuda, math=math, tol=5e-5, gtol=3e-3)


# ------------------------------------------------------------------------------------------ tcgen05
TC_TOL = {'tf32x3': dict(tol=2e-5, gtol=2e-4), 'tf32': dict(tol=5e-3, gtol=2e-2),
          'f16x3': dict(tol=2e-5, gtol=2e-4)}     # 3-term fp16 (power-of-two scales): the fp32-level bound of tf32x3


def _tc_count():
    from dl4ds_b200 import _lib
    return _lib.load().dl4ds_tc_launch_count()


@pytest.mark.parametrize('math', ['tf32x3', 'tf32', 'f16x3'])
@pytest.mark.parametrize('shape,cout,k,act', [
    ((2, 16, 16, 8), 16, 3, 'relu'),        # SW32 chunks, BW=16 BH=8
    ((2, 8, 16, 24), 40, 3, None),          # Cin 24 -> three 8-channel chunks; Cout 40 -> Npad 48
    ((1, 32, 32, 48), 48, 3, 'tanh'),       # SW64 chunks
    ((1, 64, 64, 16), 48, 1, 'relu'),       # 1x1
    ((1, 128, 128, 16), 8, 3, None),        # one row per tile, Npad 16 > Cout 8
    ((2, 16, 8, 32), 64, 5, 'sigmoid'),     # SW128 chunks, 5x5, BW=8 BH=16
    ((1, 2, 256, 8), 8, 3, 'relu'),         # W > 128: two tiles per row
])
def test_conv_tc(cuda, math, shape, cout, k, act):
    """tcgen05 implicit-GEMM forward + dgrad against the oracle; asserts the tensor-core kernel ran."""
    if math == 'tf32' and act == 'relu':
        act = 'tanh'    # single-pass tf32 flips relu masks where |y| ~ 1e-3: compare smooth graphs only
    fn = lambda c, xs: c.conv(xs[0], 'cv', cout, k=k, act=act)
    ofn = _o(lambda p, xs: R.act(R._conv(p, 'cv', xs[0], cout, k=k), act))
    n0 = _tc_count()
    compare(fn, ofn, [shape], cuda, math=math, **TC_TOL[math])
    assert _tc_count() >= n0 + 2, 'tensor-core path did not run (fwd + dgrad expected)'


@pytest.mark.parametrize('math', ['tf32x3', 'tf32', 'f16x3'])
def test_conv_tc_residual_and_d2s(cuda, math):
    a = 'relu' if math != 'tf32' else 'tanh'

    def fn(c, xs):
        y = c.conv(xs[0], 'cv', 16, act=a, res=xs[1])
        return c.conv(y, 'up', 64, d2s=2)
    ofn = _o(lambda p, xs: R.depth_to_space(R._conv(p, 'up', R.act(R._conv(p, 'cv', xs[0], 16) + xs[1], a), 64), 2))
    n0 = _tc_count()
    compare(fn, ofn, [(2, 16, 32, 8), (2, 16, 32, 16)], cuda, math=math, **TC_TOL[math])
    assert _tc_count() >= n0 + 4


@pytest.mark.parametrize('math', ['tf32x3', 'tf32', 'f16x3'])
def test_spc_block_tc(cuda, math):
    """The headline layer: shared 48 -> 192 3x3 conv + depth_to_space applied at 32^2 and 64^2."""
    fn = lambda c, xs: B.subpixel_block(c, 'spc', xs[0], 4, 48)
    ofn = _o(lambda p, xs: R.subpixel_block(p, 'spc', xs[0], 4, 48))
    compare(fn, ofn, [(1, 32, 32, 48)], cuda, math=math, **TC_TOL[math])


@pytest.mark.parametrize('math', ['tf32x3', 'f16x3'])
def test_net_resnet_spc_tc(cuda, math):
    m = nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (32, 32), n_blocks=3)
    ofn = lambda p, xs: R.net_postupsampling(p, xs, 'resnet', 'spc', 4, n_blocks=3)
    n0 = _tc_count()
    # gtol: a relu mask that flips on a ~1e-6 pre-activation difference moves a 2048-pixel weight
    # gradient by ~1e-3 of its max; the per-op tests above (smooth graphs) hold 2e-4
    compare(m.fn, ofn, [(2, 32, 32, 1)], cuda, math=math, tol=5e-5, gtol=3e-3, input_grads=False)
    assert _tc_count() > n0 + 10


# ------------------------------------------------------------------------------------------ narrow layers
@pytest.mark.parametrize('cin,cout,k', [(8, 8, 3), (8, 1, 3), (1, 8, 3), (1, 1, 3),
                                        (8, 1, 7), (1, 8, 7), (1, 1, 7),       # 7x7: the ConvNeXt stem / tail
                                        (2, 8, 3), (4, 8, 3), (8, 4, 3),       # cfg5's first layer (HR field + static variable), cfg4's 4-channel tail
                                        (1, 48, 3), (1, 16, 3), (1, 64, 3)])   # first layer of the auxiliary branch (cfg3: 1 -> 48)
@pytest.mark.parametrize('hw', [(128, 128), (64, 32), (16, 256)])
def test_thin_wgrad(cuda, cin, cout, k, hw):
    """Direct conv (fwd + dgrad) and sliding-window wgrad of the HR-tail / stem layers (thin.cu)."""
    fn = lambda c, xs: c.conv(xs[0], 'cv', cout, k=k, act='tanh')
    ofn = _o(lambda p, xs: R.act(R._conv(p, 'cv', xs[0], cout, k=k), 'tanh'))
    n = max(1, 16384 // (hw[0] * hw[1])) + 1
    compare(fn, ofn, [(n, hw[0], hw[1], cin)], cuda)


@pytest.mark.parametrize('math', ['tf32x3', 'tf32', 'f16x3'])
@pytest.mark.parametrize('hw', [(128, 128), (64, 32), (16, 256), (72, 64)])
def test_thin_mma_8x8(cuda, math, hw):
    """The 8 -> 8 3x3 layers of the HR tail in the tensor-core math modes (thin_mma.cu: mma.sync kernels; the 3-term
    mode splits into fp16 hi / lo halves under per-tile power-of-two scales): forward, dgrad, wgrad, a residual +
    relu epilogue, and inputs far from fp16's range."""
    n = max(1, 16384 // (hw[0] * hw[1])) + 1
    fn = lambda c, xs: c.conv(c.conv(xs[0], 'a', 8, k=3, act='tanh'), 'b', 8, k=3, act='relu', res=xs[1])
    ofn = _o(lambda p, xs: R.act(R._conv(p, 'b', R.act(R._conv(p, 'a', xs[0], 8, k=3), 'tanh'), 8, k=3) + xs[1], 'relu'))
    if math == 'tf32':      # single-pass mode: the smooth layer only (a relu mask flip moves the small bias gradients by percents)
        fn = lambda c, xs: c.conv(xs[0], 'a', 8, k=3, act='tanh')
        ofn = _o(lambda p, xs: R.act(R._conv(p, 'a', xs[0], 8, k=3), 'tanh'))
    compare(fn, ofn, [(n, hw[0], hw[1], 8), (n, hw[0], hw[1], 8)][:1 if math == 'tf32' else 2], cuda, math=math, **TC_TOL[math])
    if math != 'tf32':
        fn2 = lambda c, xs: c.conv(xs[0], 'cv', 8, k=3)
        ofn2 = _o(lambda p, xs: R._conv(p, 'cv', xs[0], 8, k=3))
        for sc in (3000.0, 1e-6):
            compare(fn2, ofn2, [(n, hw[0], hw[1], 8)], cuda, math=math, scale_inputs=sc, **TC_TOL[math])


def test_bias_act_bwd_vec4_d2s(cuda):
    fn = lambda cx, xs: cx.conv(xs[0], 'cv', 48 * 4, d2s=2)
    ofn = _o(lambda p, xs: R.depth_to_space(R._conv(p, 'cv', xs[0], 48 * 4), 2))
    compare(fn, ofn, [(2, 8, 8, 16)], cuda)


def test_thin_conv_residual_bias_relu(cuda):
    def fn(c, xs):
        return c.conv(xs[0], 'cv', 8, act='relu', res=xs[1])
    ofn = _o(lambda p, xs: R.act(R._conv(p, 'cv', xs[0], 8) + xs[1], 'relu'))
    compare(fn, ofn, [(2, 96, 128, 8), (2, 96, 128, 8)], cuda)


@pytest.mark.parametrize('cin,cout', [(48, 8), (8, 48), (8, 8), (16, 8)])
def test_pointwise_conv(cuda, cin, cout):
    """Streaming 1x1 kernel of the narrow layers (TransitionLast and its dgrad), bias + relu + residual."""
    def fn(c, xs):
        return c.conv(xs[0], 'cv', cout, k=1, act='relu', res=xs[1])
    ofn = _o(lambda p, xs: R.act(R._conv(p, 'cv', xs[0], cout, k=1) + xs[1], 'relu'))
    compare(fn, ofn, [(5, 128, 128, cin), (5, 128, 128, cout)], cuda)


def test_pointwise_conv_channel_tails(cuda):
    """cfg3's tail wiring: a 48 -> 2 1x1 convolution (LocalizedConvBlock transition), the nested concatenation
    [[48 | 2] | 48] = 98 channels (padded pitch 100, the last part starts at channel 50: its gradient slice is
    re-aligned) and the 98 -> 8 1x1 convolution (TransitionLast) -- the streaming 1x1 kernels with channel tails."""
    def fn(c, xs):
        lc = c.conv(xs[0], 'lc', 2, k=1, act='tanh')
        aux = c.conv(xs[1], 'aux', 48, k=3, act='tanh')
        cat = c.concat([c.concat([xs[0], lc]), aux])
        return c.conv(cat, 'tl', 8, k=1, act='tanh')

    def ofn(p, xs):
        x0, x1 = R._nchw(xs[0]), R._nchw(xs[1])
        lc = R.act(R._conv(p, 'lc', x0, 2, k=1), 'tanh')
        aux = R.act(R._conv(p, 'aux', x1, 48, k=3), 'tanh')
        cat = torch.cat([torch.cat([x0, lc], 1), aux], 1)
        return R._nhwc(R.act(R._conv(p, 'tl', cat, 8, k=1), 'tanh'))
    for math in ('fp32', 'tf32x3'):
        compare(fn, ofn, [(5, 128, 128, 48), (5, 128, 128, 8)], cuda, math=math, **TC_TOL['tf32x3'])


# ------------------------------------------------------------------------------------------ stacked-taps wgrad
@pytest.mark.parametrize('math', ['tf32x3', 'tf32', 'f16x3'])
@pytest.mark.parametrize('shape,cout,k', [
    ((3, 32, 32, 48), 48, 3),       # backbone layer: 432 stacked rows -> 4 M-blocks, one role
    ((2, 64, 64, 48), 192, 3),      # SPC layer: two output-channel roles of 96, two chunks per image row
    ((2, 16, 16, 24), 8, 3),        # 2 rows per chunk (BW=16), 216 rows -> 2 blocks, Nmma 16 > Cb 8
    ((2, 8, 8, 64), 16, 3),         # 4 rows per chunk (BW=8); 576 rows -> two input-channel groups of 32
    ((1, 128, 128, 48), 8, 1),      # TransitionLast: 1x1, long pixel stream, narrow N
    ((2, 32, 32, 8), 48, 1),        # 1x1 projection, 8 stacked rows
    ((2, 16, 32, 16), 40, 5),       # 5x5: 400 stacked rows, halo 36 x 5
    ((5, 32, 32, 40), 56, 3),       # 360 rows (3 blocks, last one partial), Cb 56 -> Nmma 64
    ((2, 32, 32, 8), 8, 7),         # 7x7 (ConvNeXt stem / tail): 392 stacked rows, halo 38 x 7
    ((1, 64, 128, 8), 16, 7),
])
def test_wgrad_stacked_taps(cuda, math, shape, cout, k):
    """conv_tc_wgrad2_kernel (conv_tc_wgrad.cu) through Ctx.conv's backward, against the oracle."""
    fn = lambda c, xs: c.conv(xs[0], 'cv', cout, k=k, act='tanh')
    ofn = _o(lambda p, xs: R.act(R._conv(p, 'cv', xs[0], cout, k=k), 'tanh'))
    compare(fn, ofn, [shape], cuda, math=math, **TC_TOL[math])


@pytest.mark.parametrize('shape,cout,k', [
    ((3, 32, 32, 48), 48, 3),       # backbone layer: one chunk per image row
    ((2, 32, 64, 32), 32, 3),       # two chunks per row: the halo columns come from the neighbouring chunk
    ((1, 64, 128, 40), 32, 3),      # four chunks per row, 24 padding channels per panel
    ((2, 32, 32, 64), 48, 3),       # full 64-channel panels
    ((2, 32, 32, 16), 16, 3),       # narrow layer: below the width threshold, conv_tc_wgrad2 serves it
    ((70, 32, 32, 40), 56, 3),      # 2240 chunks: 16 per CTA, the stage ring wraps; Cb 56 -> 64 columns x 9 taps > 512: wgrad2
    ((70, 32, 32, 48), 40, 3),      # the same through this kernel: range maxima over many images
])
def test_wgrad_mn_major_f16(cuda, shape, cout, k):
    """conv_tc_wgrad3_kernel (kind::f16, MN-major operands read from the NHWC tiles, per-CTA power-of-two scales)
    through Ctx.conv's backward against the oracle, at the 3xTF32 bounds; asserts the kernel's domain covers the case."""
    from dl4ds_b200 import _lib
    fn = lambda c, xs: c.conv(xs[0], 'cv', cout, k=k, act='tanh')
    ofn = _o(lambda p, xs: R.act(R._conv(p, 'cv', xs[0], cout, k=k), 'tanh'))
    compare(fn, ofn, [shape], cuda, math='tf32x3', **TC_TOL['tf32x3'])
    # activations far outside fp16's comfortable range (what the scales are for); linear layer: no tanh saturation
    fn2 = lambda c, xs: c.conv(xs[0], 'cv', cout, k=k)
    ofn2 = _o(lambda p, xs: R._conv(p, 'cv', xs[0], cout, k=k))
    compare(fn2, ofn2, [shape], cuda, math='tf32x3', scale_inputs=3000.0, **TC_TOL['tf32x3'])
    compare(fn2, ofn2, [shape], cuda, math='tf32x3', scale_inputs=1e-6, **TC_TOL['tf32x3'])


def test_wgrad_stacked_taps_on_concat_slices(cuda):
    """Q (dz) as a channel slice of a wider gradient buffer (pitch 24, 16 channels): the gradient of a
    Concatenate input is a slice of the concat's gradient (dense-block wiring, blocks.py:276)."""
    def fn(c, xs):
        a = c.conv(xs[0], 'a', 16, act='tanh')
        cat = c.concat([a, xs[0]])                       # 16 + 8 channels
        return c.conv(cat, 'b', 24, act='tanh')

    def ofn(p, xs):
        x = R._nchw(xs[0])
        a = R.act(R._conv(p, 'a', x, 16), 'tanh')
        return R._nhwc(R.act(R._conv(p, 'b', torch.cat([a, x], 1), 24), 'tanh'))
    compare(fn, ofn, [(2, 32, 32, 8)], cuda, math='tf32x3', **TC_TOL['tf32x3'])


# ------------------------------------------------------------------------------------------ composed SPC stage + 1x1
@pytest.mark.parametrize('math', ['fp32', 'tf32x3', 'f16x3'])
@pytest.mark.parametrize('scale,cin,hw', [(4, 48, 16), (2, 16, 32), (4, 8, 32)])
def test_subpixel_transition_composed(cuda, math, scale, cin, hw):
    """SubpixelConvolutionBlock + TransitionLast with the last x2 stage composed with the 1x1 convolution
    (Ctx.conv_d2s_pointwise) against the oracle's two separate blocks: forward, input gradient and all four
    parameter gradients (chain rule back onto conv2x / TransitionLast, shared conv2x summed over stages)."""
    fn = lambda c, xs: B.subpixel_transition(c, 'spc', xs[0], scale, cin, 'tl', 8, 'tanh')
    ofn = _o(lambda p, xs: R.transition_block(p, 'tl', R.subpixel_block(p, 'spc', xs[0], scale, cin), 8, 'tanh'))
    tol = dict(tol=2e-5, gtol=2e-4) if math == 'fp32' else TC_TOL[math]
    compare(fn, ofn, [(2, hw, hw, cin)], cuda, math=math, **tol)


def test_subpixel_transition_fallback_x5(cuda):
    """scale 10 = x2 then x5: the last stage is not x2, the two blocks run unfused."""
    fn = lambda c, xs: B.subpixel_transition(c, 'spc', xs[0], 10, 4, 'tl', 8, 'relu')
    ofn = _o(lambda p, xs: R.transition_block(p, 'tl', R.subpixel_block(p, 'spc', xs[0], 10, 4), 8, 'relu'))
    compare(fn, ofn, [(1, 6, 6, 4)], cuda)


# ------------------------------------------------------------------------------------------ Conv2DTranspose on tensor cores
@pytest.mark.parametrize('math', ['fp32', 'tf32x3', 'f16x3'])
@pytest.mark.parametrize('shape,cout,k,stride,act', [
    ((2, 16, 16, 8), 48, 9, 2, None),        # DeconvolutionBlock T1 (blocks.py:508-516): 8 -> 48, 9x9, s=2
    ((2, 32, 32, 48), 48, 9, 2, 'tanh'),     # T2: 48 -> 48 with the block's activation
    ((1, 16, 16, 16), 8, 5, 2, None),        # another odd kernel
    ((1, 8, 8, 8), 8, 9, 4, None),           # stride 4 (scale-4 fall-through): depth_to_space(4) -> CUDA-core store path
])
def test_conv_transpose_as_conv_d2s(cuda, math, shape, cout, k, stride, act):
    """Ctx.conv_transpose = rearranged stride-1 convolution + depth_to_space (dl4ds_convt_rearrange) against the
    oracle's conv_transpose: forward, input gradient and the gradient scattered back onto the Keras kernel."""
    fn = lambda c, xs: c.conv_transpose(xs[0], 'ct', cout, k, stride, act=act)

    def ofn_(p, xs):
        w = p.get('ct/kernel', (k, k, cout, xs[0].shape[1]))
        return R.act(R.conv2d_transpose_same(xs[0], w, stride), act)
    tol = dict(tol=2e-5, gtol=2e-4) if math == 'fp32' else TC_TOL[math]
    n0 = _tc_count()
    compare(fn, _o(ofn_), [shape], cuda, math=math, **tol)
    if math != 'fp32' and stride == 2 and shape[2] >= 16:
        assert _tc_count() > n0, 'tensor-core path did not run'


@pytest.mark.parametrize('shape,cout,act,with_res', [((4, 8, 8, 256), 256, 'relu', False),
                                                      ((4, 4, 4, 256), 256, 'tanh', True),
                                                      ((3, 8, 8, 128), 64, None, False)])
def test_conv_small_map_split_k(cuda, shape, cout, act, with_res):
    """Deep U-Net levels (4x4 / 8x8 maps, 128-256 channels): the CUDA-core kernel splits the (tap, channel) loop over
    blockIdx.z and finishes bias / residual / activation in a second pass (conv_finish_kernel)."""
    if with_res:
        fn = lambda c, xs: c.conv(xs[0], 'cv', cout, k=3, act=act, res=xs[1])
        ofn = _o(lambda p, xs: R.act(R._conv(p, 'cv', xs[0], cout, k=3) + xs[1], act))
        shapes = [shape, shape[:3] + (cout,)]
    else:
        fn = lambda c, xs: c.conv(xs[0], 'cv', cout, k=3, act=act)
        ofn = _o(lambda p, xs: R.act(R._conv(p, 'cv', xs[0], cout, k=3), act))
        shapes = [shape]
    compare(fn, ofn, shapes, cuda, math='tf32x3', tol=2e-5, gtol=2e-4)


def test_net_recnet_pin(cuda):
    """recnet_pin (spt_preups.py:12-163): ConvLSTM backbone on pre-upsampled samples, static HR branch, T=3."""
    T, Bz = 3, 2
    m = nets.recnet_pin('resnet', 2, 1, (16, 16), T, n_blocks=1)
    assert m.name == 'recresnet_pin'

    def ofn(p, xs):
        x = xs[0]
        x5 = x.reshape(T, Bz, *x.shape[1:]).permute(1, 0, 2, 3, 4)      # (B,T,H,W,C)
        y5 = R.recnet_pin(p, [x5, xs[1]], 'resnet', T, n_blocks=1)
        return y5.permute(1, 0, 2, 3, 4).reshape(T * Bz, *y5.shape[2:])
    shapes = [(T * Bz, 16, 16, 2), (Bz, 16, 16, 1)]
    compare(m.fn, ofn, shapes, cuda, tol=5e-5, gtol=3e-3, input_grads=False)
