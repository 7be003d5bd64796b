"""Oracle self-consistency (CPU): the torch-fp32 restatement (oracle/torch_ref.py) against the
independent numpy-fp64 direct-loop definitions (oracle/ops_np.py), and the structural known-answers
the reference's notebook records (ipynb:3335-3383: per-layer parameter counts, total 487 511)."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from dl4ds_b200 import nets
from oracle import ops_np as N
from oracle import torch_ref as R

RNG = np.random.default_rng(11)


def _t(a):   # NHWC numpy -> NCHW torch
    return torch.as_tensor(a).permute(0, 3, 1, 2).contiguous()


def _n(t):   # NCHW torch -> NHWC numpy
    return t.permute(0, 2, 3, 1).numpy()


def _close(a, b, tol=1e-5):
    assert np.abs(np.asarray(a, np.float64) - b).max() <= tol * max(1.0, np.abs(b).max())


@pytest.mark.parametrize('h,w,k,s,pad', [(7, 6, 3, 1, 'same'), (8, 8, 3, 2, 'same'), (9, 7, 3, 2, 'same'),
                                         (9, 9, 3, 2, 'valid'), (6, 6, 5, 1, 'same'), (5, 5, 1, 1, 'same')])
def test_conv2d(h, w, k, s, pad):
    x = RNG.standard_normal((2, h, w, 3)).astype(np.float32)
    wt = RNG.standard_normal((k, k, 3, 4)).astype(np.float32)
    b = RNG.standard_normal(4).astype(np.float32)
    y = R.conv2d(_t(x), torch.as_tensor(wt), torch.as_tensor(b), stride=s, padding=pad)
    _close(_n(y), N.conv2d(x, wt, b, stride=s, padding=pad))


@pytest.mark.parametrize('s', [2, 4])
def test_conv2d_transpose(s):
    x = RNG.standard_normal((2, 4, 3, 3)).astype(np.float32)
    wt = RNG.standard_normal((9, 9, 5, 3)).astype(np.float32)
    y = R.conv2d_transpose_same(_t(x), torch.as_tensor(wt), s)
    ref = N.conv2d_transpose_same(x, wt, s)
    assert ref.shape == (2, 4 * s, 3 * s, 5)
    _close(_n(y), ref, 2e-5)


def test_conv2d_transpose_is_same_conv_adjoint():
    """Conv2DTranspose(padding='same') is defined as the input-gradient of the SAME conv."""
    x = torch.randn(1, 3, 6, 6, dtype=torch.float64, requires_grad=True)
    w = torch.randn(9, 9, 3, 4, dtype=torch.float64)       # as Conv2D HWIO (3 -> 4)
    y = R.conv2d(x, w, None, stride=2, padding='same')
    g = torch.randn_like(y)
    y.backward(g)
    # the same kernel read as Conv2DTranspose (kh,kw,Cout=3,Cin=4)
    up = R.conv2d_transpose_same(g, w, 2)
    assert torch.allclose(up, x.grad, atol=1e-10)


@pytest.mark.parametrize('r', [2, 3, 5])
def test_depth_to_space(r):
    x = RNG.standard_normal((2, 3, 4, 2 * r * r)).astype(np.float32)
    assert np.array_equal(_n(R.depth_to_space(_t(x), r)), N.depth_to_space(x, r))


def test_depth_to_space_is_not_pixel_shuffle():
    x = torch.randn(1, 8, 2, 2)
    assert not torch.equal(R.depth_to_space(x, 2), torch.nn.functional.pixel_shuffle(x, 2))


@pytest.mark.parametrize('ho,wo', [(12, 20), (3, 2), (7, 9)])
def test_resize_bilinear(ho, wo):
    x = RNG.standard_normal((2, 6, 5, 2)).astype(np.float32)
    _close(_n(R.resize_bilinear(_t(x), ho, wo)), N.resize_bilinear(x, ho, wo))


def test_maxpool_local_attention():
    x = RNG.standard_normal((2, 7, 6, 4)).astype(np.float32)
    assert np.array_equal(_n(R.maxpool2(_t(x))), N.maxpool2(x))
    w = RNG.standard_normal((7, 6, 4, 2)).astype(np.float32)
    b = RNG.standard_normal((7, 6, 2)).astype(np.float32)
    _close(_n(R.local_conv1x1(_t(x), torch.as_tensor(w), torch.as_tensor(b))), N.local_conv1x1(x, w, b))
    ws = OrderedDict([('a/conv1/kernel', RNG.standard_normal((1, 1, 4, 1)).astype(np.float32)),
                      ('a/conv1/bias', RNG.standard_normal(1).astype(np.float32)),
                      ('a/conv2/kernel', RNG.standard_normal((1, 1, 1, 4)).astype(np.float32)),
                      ('a/conv2/bias', RNG.standard_normal(4).astype(np.float32))])
    p = R.Params({k: torch.as_tensor(v) for k, v in ws.items()})
    y = R.channel_attention(p, 'a', _t(x), 4)
    _close(_n(y), N.channel_attention(x, *ws.values()))


def test_convlstm():
    x = RNG.standard_normal((2, 3, 5, 4, 2)).astype(np.float32)       # (B,T,H,W,C)
    wx = (0.3 * RNG.standard_normal((3, 3, 2, 12))).astype(np.float32)
    wh = (0.3 * RNG.standard_normal((3, 3, 3, 12))).astype(np.float32)
    b = RNG.standard_normal(12).astype(np.float32)
    p = R.Params({'l/kernel': torch.as_tensor(wx), 'l/recurrent_kernel': torch.as_tensor(wh),
                  'l/bias': torch.as_tensor(b)})
    y = R.convlstm2d(p, 'l', torch.as_tensor(x).permute(0, 1, 4, 2, 3), 3, 3)
    _close(y.permute(0, 1, 3, 4, 2).numpy(), N.convlstm2d(x, wx, wh, b), 2e-5)


def test_bce_adam_blockmean():
    pr = (0.01 + 0.98 * RNG.random(16)).astype(np.float32)   # fp32 vs fp64 differ at the eps clip edges
    for t in (0.0, 1.0):
        _close(float(R.bce(torch.full((16,), t), torch.as_tensor(pr))), N.bce(t, pr.astype(np.float64)), 1e-5)
    th = RNG.standard_normal(50).astype(np.float32)
    w = {'w': torch.as_tensor(th.copy())}
    opt = R.TFAdam(['w'], lr=1e-2)
    th64, m, v = th.astype(np.float64), 0.0, 0.0
    for t in range(1, 4):
        g = RNG.standard_normal(50).astype(np.float32)
        opt.apply(w, {'w': torch.as_tensor(g)})
        th64, m, v = N.adam_step(th64, g.astype(np.float64), m, v, t, 1e-2)
    _close(w['w'].numpy(), th64, 1e-5)
    x = RNG.standard_normal((1, 8, 8, 1))
    assert N.block_mean(x, 4).shape == (1, 2, 2, 1)


def test_piecewise_constant_decay():
    f = R.piecewise_constant(10, 1e-3, 1e-4)
    assert f(10) == 1e-3 and f(11) == 1e-4


# ---------------------------------------------------------------------------------- structure
NOTEBOOK_TABLE = [152, 1168, 3632, 9096, 16992, 27320, 40080, 55272, 72896, 36928, 576, 147712, 73858,
                  536, 1210, 83]   # notebooks/DL4DS_tutorial.ipynb:3335-3383, total 487 511


def _oracle_spec(fn, shapes):
    p = R.Params()
    fn(p, [torch.zeros(s) for s in shapes])
    return p.spec


def test_notebook_parameter_table():
    spec = _oracle_spec(lambda p, xs: R.net_postupsampling(p, xs, 'resnet', 'spc', 8, n_blocks=8,
                                                           localcon_layer=True), [(1, 12, 16, 2)])
    groups = OrderedDict()
    for n, s in spec.items():
        groups[n.split('/')[0]] = groups.get(n.split('/')[0], 0) + int(np.prod(s))
    assert sum(groups.values()) == 487511
    assert sorted(groups.values()) == sorted(NOTEBOOK_TABLE)
    m = nets.net_postupsampling('resnet', 'spc', 8, 2, 0, (12, 16), n_blocks=8, localcon_layer=True)
    assert m.count_params() == 487511 and m.output_shape == (96, 128, 1)
    assert dict(m.spec) == dict(spec)


CASES = [
    ('cfg1', lambda: nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (32, 32)),
     lambda p, xs: R.net_postupsampling(p, xs, 'resnet', 'spc', 4), [(1, 32, 32, 1)], 204405, 573005856),
    ('cfg3', lambda: nets.net_postupsampling('densenet', 'dc', 8, 5, 1, (16, 16), attention=True,
                                             localcon_layer=True),
     lambda p, xs: R.net_postupsampling(p, xs, 'densenet', 'dc', 8, attention=True, localcon_layer=True),
     [(1, 16, 16, 5), (1, 128, 128, 1)], 596026, None),
    ('cfg5G', lambda: nets.unet_pin('unet', 2, 1, (256, 256), 1, 8, 6),
     lambda p, xs: R.unet_pin(p, xs, 8, 6), [(1, 256, 256, 2), (1, 256, 256, 1)], 5705365, None),
    ('cfg5D', lambda: nets.residual_discriminator(2, 'pin', False, 4, (64, 64)),
     lambda p, xs: R.residual_discriminator(p, xs, 'pin', 4, (64, 64)), [(1, 64, 64, 2), (1, 64, 64, 1)],
     15961, None),
    ('pin', lambda: nets.net_pin('densenet', 3, 0, (16, 16), n_blocks=2),
     lambda p, xs: R.net_pin(p, xs, 'densenet', n_blocks=2), [(1, 16, 16, 3)], None, None),
]


@pytest.mark.parametrize('name,build,ofn,shapes,nparams,macs', CASES, ids=[c[0] for c in CASES])
def test_product_graph_matches_oracle_structure(name, build, ofn, shapes, nparams, macs):
    m = build()
    spec = _oracle_spec(ofn, shapes)
    assert dict(m.spec) == dict(spec)
    if nparams is not None:
        assert m.count_params() == nparams
    if macs is not None:
        assert m.macs_per_sample == macs


def test_recnet_structure_cfg4():
    m = nets.recnet_postupsampling('resnet', 'rc', 4, 1, 0, (32, 32), 6)
    p = R.Params()
    R.recnet_postupsampling(p, [torch.zeros(1, 6, 32, 32, 1)], 'resnet', 'rc', 4, 6)
    assert dict(m.spec) == dict(p.spec) and m.count_params() == 83385
    assert m.name == 'recresnet_rc' and len(m.input.shape) == 5


def test_model_names_and_unsupported():
    assert nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (8, 8)).name == 'resnet_spc'
    assert nets.unet_pin('unet', 1, 0, (16, 16), 1, 8, 2).name == 'unet_pin'
    with pytest.raises(ValueError):
        nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (8, 8), normalization='gn')
    with pytest.raises(ValueError):
        nets.recnet_postupsampling('resnet', 'spc', 4, 1, 0, (8, 8), 3, normalization='gn')
    with pytest.raises(ValueError):           # ConvNextBlock without a normalisation fails in the reference too
        nets.net_postupsampling('convnext', 'spc', 4, 1, 0, (8, 8))
    with pytest.raises(NotImplementedError):
        nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (8, 8), activation='elu')
    with pytest.raises(ValueError):
        nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (8, 8), dropout_rate=0.2, dropout_variant='alpha')
    with pytest.raises(ValueError):
        nets.recnet_postupsampling('resnet', 'spc', 4, 1, 0, (8, 8), 3, dropout_rate=0.2, dropout_variant='alpha')


def test_recnet_pin_structure():
    """recnet_pin: same parameters as recnet_postupsampling minus the upsampler; TransitionLast -> n_filters
    (spt_preups.py:133)."""
    from dl4ds_b200 import nets
    m = nets.recnet_pin('resnet', 1, 0, (16, 16), 4, n_filters=8, n_blocks=2)
    assert m.name == 'recresnet_pin' and m.input.shape == (None, 4, 16, 16, 1)
    assert m.spec['TransitionLast/conv/kernel'] == (1, 1, 8, 8)
    assert not any(k.startswith(('SubpixelConvolution', 'ResizeConvolution', 'Deconvolution')) for k in m.spec)
    md = nets.recnet_pin('densenet', 1, 0, (16, 16), 4, n_filters=8, n_blocks=2)
    assert md.spec['TransitionLast/conv/kernel'] == (1, 1, 16, 8)
