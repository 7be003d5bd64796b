"""CPU tests of the ConvNeXt pieces (blocks.py:131-184): the oracle's depthwise convolution and GELU against direct
numpy fp64 statements, the block's parameter table by hand, builder / oracle agreement for the 'convnext' backbone."""
import numpy as np
import pytest
import torch
from scipy.special import erf

from dl4ds_b200 import nets, utils
from oracle import torch_ref as R


def test_depthwise_conv_against_direct_loops():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 3, 6, 5))                  # NCHW
    w = rng.standard_normal((7, 7, 3, 1))
    b = rng.standard_normal(3)
    y = R.depthwise_conv2d(torch.tensor(x), torch.tensor(w), torch.tensor(b)).numpy()
    ref = np.zeros_like(x)
    for n in range(2):
        for c in range(3):
            for h in range(6):
                for q in range(5):
                    acc = b[c]
                    for i in range(7):
                        for j in range(7):
                            hh, ww = h + i - 3, q + j - 3
                            if 0 <= hh < 6 and 0 <= ww < 5:
                                acc += w[i, j, c, 0] * x[n, c, hh, ww]
                    ref[n, c, h, q] = acc
    assert np.abs(y - ref).max() < 1e-12


def test_gelu_is_the_exact_erf_form():
    v = np.linspace(-6, 6, 101)
    assert np.abs(R.act(torch.tensor(v), 'gelu').numpy() - 0.5 * v * (1 + erf(v / np.sqrt(2)))).max() < 1e-12


def test_convnext_block_parameter_table_by_hand():
    p = R.Params()
    R.convnext_block(p, 'b', torch.zeros(1, 16, 8, 8), 24, 'gelu', 'ln', use_1x1conv=True)
    assert dict(p.spec) == {
        'b/dwconv/depthwise_kernel': (7, 7, 16, 1), 'b/dwconv/bias': (16,),
        'b/norm/gamma': (16,), 'b/norm/beta': (16,),
        'b/pwconv1/kernel': (16, 96), 'b/pwconv1/bias': (96,),
        'b/pwconv2/kernel': (96, 24), 'b/pwconv2/bias': (24,),
        'b/conv1x1/kernel': (1, 1, 16, 24), 'b/conv1x1/bias': (24,)}
    with pytest.raises(ValueError):          # the reference has no `self.norm` then (blocks.py:158-164,173)
        R.convnext_block(R.Params(), 'b', torch.zeros(1, 16, 8, 8), 16, 'gelu', None)


@pytest.mark.parametrize('nz', ['ln', 'bn'])
def test_convnext_builders_match_oracle(nz):
    m = nets.net_postupsampling('convnext', 'spc', 4, 1, 1, (16, 16), n_blocks=3, normalization=nz, activation='gelu')
    p = R.Params()
    y = R.net_postupsampling(p, [torch.zeros(1, 16, 16, 1), torch.zeros(1, 64, 64, 1)], 'convnext', 'spc', 4,
                             n_blocks=3, normalization=nz, activation='gelu')
    assert dict(m.spec) == dict(p.spec) and tuple(y.shape) == (1, 64, 64, 1) and m.name == 'convnext_spc'
    assert m.spec['stem/kernel'] == (7, 7, 1, 8) and m.spec['ConvBlock_out/conv2/kernel'][:2] == (7, 7)
    m = nets.net_pin('convnext', 2, 0, (16, 16), n_blocks=2, normalization=nz)
    p = R.Params()
    R.net_pin(p, [torch.zeros(1, 16, 16, 2)], 'convnext', n_blocks=2, normalization=nz)
    assert dict(m.spec) == dict(p.spec)


def test_convnext_needs_a_normalisation_and_spatial_samples():
    with pytest.raises(ValueError):
        nets.net_postupsampling('convnext', 'spc', 4, 1, 0, (16, 16))
    with pytest.raises(ValueError):
        utils.check_compatibility_upsbackb('convnext', 'spc', 4)
