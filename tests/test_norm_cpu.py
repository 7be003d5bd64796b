"""CPU tests of the normalisation layers (blocks.py:63-71): the oracle's BatchNormalization / LayerNormalization
restatement against direct numpy fp64 formulas, the moving-statistics rule, and parameter-table agreement between
the product's graph builder and the oracle for networks built with ``normalization='bn' | 'ln'``."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from dl4ds_b200 import nets
from oracle import torch_ref as R


def _weights(c, rng, bn):
    w = OrderedDict()
    w['n/gamma'] = torch.tensor(1 + 0.1 * rng.standard_normal(c))
    w['n/beta'] = torch.tensor(0.1 * rng.standard_normal(c))
    if bn:
        w['n/moving_mean'] = torch.tensor(0.1 * rng.standard_normal(c))
        w['n/moving_variance'] = torch.tensor(1 + 0.1 * rng.random(c))
    return w


def test_layernorm_formula():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 5, 4, 7))                       # NCHW
    w = _weights(5, rng, False)
    y = R.normalize(R.Params(w, dtype=torch.float64), 'n', torch.tensor(x), 'ln').numpy()
    mu = x.mean(axis=1, keepdims=True)
    var = x.var(axis=1, keepdims=True)                          # biased
    ref = (x - mu) / np.sqrt(var + 1e-3) * w['n/gamma'].numpy().reshape(1, 5, 1, 1) + w['n/beta'].numpy().reshape(1, 5, 1, 1)
    assert np.abs(y - ref).max() < 1e-12


def test_batchnorm_training_inference_and_moving_statistics():
    rng = np.random.default_rng(1)
    x = 2 + 3 * rng.standard_normal((4, 3, 6, 5))
    w = _weights(3, rng, True)
    mm0, mv0 = w['n/moving_mean'].numpy().copy(), w['n/moving_variance'].numpy().copy()
    g, b = w['n/gamma'].numpy().reshape(1, 3, 1, 1), w['n/beta'].numpy().reshape(1, 3, 1, 1)
    p = R.Params(w, dtype=torch.float64, training=True)
    y = R.normalize(p, 'n', torch.tensor(x), 'bn').numpy()
    mu, var = x.mean(axis=(0, 2, 3)), x.var(axis=(0, 2, 3))
    ref = (x - mu.reshape(1, 3, 1, 1)) / np.sqrt(var.reshape(1, 3, 1, 1) + 1e-3) * g + b
    assert np.abs(y - ref).max() < 1e-12
    m = x.size // 3
    assert np.allclose(w['n/moving_mean'].numpy(), 0.99 * mm0 + 0.01 * mu, atol=1e-14)
    assert np.allclose(w['n/moving_variance'].numpy(), 0.99 * mv0 + 0.01 * var * m / (m - 1), atol=1e-14)
    # inference: the moving statistics, no update
    mm1, mv1 = w['n/moving_mean'].numpy().copy(), w['n/moving_variance'].numpy().copy()
    y = R.normalize(R.Params(w, dtype=torch.float64, training=False), 'n', torch.tensor(x), 'bn').numpy()
    ref = (x - mm1.reshape(1, 3, 1, 1)) / np.sqrt(mv1.reshape(1, 3, 1, 1) + 1e-3) * g + b
    assert np.abs(y - ref).max() < 1e-12
    assert np.array_equal(w['n/moving_mean'].numpy(), mm1)
    with pytest.raises(ValueError):
        R.normalize(p, 'n', torch.tensor(x), 'gn')


CASES = [
    ('resnet', lambda nz: nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (8, 8), n_blocks=2, normalization=nz),
     lambda p, nz: R.net_postupsampling(p, [torch.zeros(1, 8, 8, 1)], 'resnet', 'spc', 4, n_blocks=2, normalization=nz)),
    ('densenet', lambda nz: nets.net_postupsampling('densenet', 'rc', 2, 2, 1, (8, 8), n_blocks=2, normalization=nz,
                                                    attention=True),
     lambda p, nz: R.net_postupsampling(p, [torch.zeros(1, 8, 8, 2), torch.zeros(1, 16, 16, 1)], 'densenet', 'rc', 2,
                                        n_blocks=2, normalization=nz, attention=True)),
    ('convnet_pin', lambda nz: nets.net_pin('convnet', 1, 0, (8, 8), n_blocks=2, normalization=nz),
     lambda p, nz: R.net_pin(p, [torch.zeros(1, 8, 8, 1)], 'convnet', n_blocks=2, normalization=nz)),
    ('unet', lambda nz: nets.unet_pin('unet', 1, 1, (16, 16), 1, 8, 2, normalization=nz),
     lambda p, nz: R.unet_pin(p, [torch.zeros(1, 16, 16, 1), torch.zeros(1, 16, 16, 1)], 8, 2, normalization=nz)),
    ('discriminator', lambda nz: nets.residual_discriminator(1, 'pin', False, 4, (16, 16), n_res_blocks=1,
                                                             normalization=nz),
     lambda p, nz: R.residual_discriminator(p, [torch.zeros(1, 16, 16, 1), torch.zeros(1, 16, 16, 1)], 'pin', 4,
                                            (16, 16), n_res_blocks=1, normalization=nz)),
]


@pytest.mark.parametrize('nz', ['bn', 'ln'])
@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_parameter_tables_with_normalization(case, nz):
    _, build, oracle = case
    m = build(nz)
    p = R.Params()
    oracle(p, nz)
    assert dict(m.spec) == dict(p.spec)
    plain = build(None)
    # a normalised conv loses its bias (C) and gains gamma + beta (2C) [+ moving mean / variance (2C)]
    assert m.count_params() > plain.count_params()
    n_norm = sum(1 for k in m.spec if k.endswith('/gamma'))
    assert n_norm > 0 and sum(1 for k in m.spec if k.endswith('/moving_mean')) == (n_norm if nz == 'bn' else 0)


def test_resnet_bn_parameter_count_by_hand():
    """ResidualBlock(8) on 8 channels with BN: 2 x (3*3*8*8) kernels, no biases, 2 x 4*8 BN variables."""
    m0 = nets.net_postupsampling('resnet', 'spc', 2, 1, 0, (8, 8), n_blocks=1)
    m1 = nets.net_postupsampling('resnet', 'spc', 2, 1, 0, (8, 8), n_blocks=1, normalization='bn')
    blk0 = sum(int(np.prod(s)) for k, s in m0.spec.items() if k.startswith('ResidualBlock1/'))
    blk1 = sum(int(np.prod(s)) for k, s in m1.spec.items() if k.startswith('ResidualBlock1/'))
    assert blk0 == 2 * (9 * 64 + 8) and blk1 == 2 * (9 * 64) + 2 * 32


@pytest.mark.parametrize('nz', ['bn', 'ln'])
def test_recurrent_networks_with_normalization_and_dropout(nz):
    """RecurrentConvBlock(normalization, dropout) and the 5-D tail -- blocks.py:339-398, spt_postups.py:104-157:
    same parameter table and the same number of dropout applications in the builder and in the oracle."""
    from dl4ds_b200.spec import SpecCtx
    kw = dict(n_blocks=1, normalization=nz, dropout_rate=0.2, dropout_variant='spatial')
    m = nets.recnet_postupsampling('resnet', 'rc', 4, 1, 1, (8, 8), 3, **kw)
    p = R.Params()
    y = R.recnet_postupsampling(p, [torch.zeros(2, 3, 8, 8, 1), torch.zeros(2, 32, 32, 1)], 'resnet', 'rc', 4, 3, **kw)
    assert dict(m.spec) == dict(p.spec) and tuple(y.shape) == (2, 3, 32, 32, 1)
    sc = SpecCtx()
    m.fn(sc, [sc.input((6, 8, 8, 1)), sc.input((2, 32, 32, 1))])
    assert sc.n_dropout == p.n_dropout == 2 + 1 + 2          # block 2, after the blocks, the tail ConvBlock
    assert 'RecurrentConvBlock1/norm1/gamma' in m.spec and 'ConvBlock_tail/conv1/bias' not in m.spec
