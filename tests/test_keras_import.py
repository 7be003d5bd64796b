"""Keras structural-name importer (dl4ds_b200/keras_import.py): names the reference-side export snippet produces
for each architecture resolve onto the parameter table; round trip export -> shuffled / renumbered file -> import."""
import numpy as np
import pytest

from dl4ds_b200 import keras_import as K
from dl4ds_b200 import nets


def _models():
    return {
        'cfg2': nets.net_postupsampling('resnet', 'spc', 4, 1, 0, (16, 16), n_blocks=2),
        'cfg3': nets.net_postupsampling('densenet', 'dc', 8, 5, 1, (8, 8), n_blocks=2, attention=True,
                                        localcon_layer=True, normalization='bn'),
        'cfg4': nets.recnet_postupsampling('resnet', 'rc', 4, 1, 1, (8, 8), 3, n_blocks=1, localcon_layer=True),
        'cfg5': nets.unet_pin('unet', 2, 1, (32, 32), 1, 4, 2),
        'convnext': nets.net_postupsampling('convnext', 'rc', 2, 1, 0, (8, 8), n_blocks=2, normalization='ln'),
        'disc': nets.residual_discriminator(2, 'pin', False, 4, (16, 16), n_res_blocks=2),
        'disc_st': nets.residual_discriminator(1, 'spc', True, 4, (8, 8), n_res_blocks=1, time_window=3),
    }


@pytest.mark.parametrize('which', list(_models()))
def test_round_trip(which):
    m = _models()[which].to('cpu').init_weights(3)
    ref = m.get_weights()
    # a Keras session in which other layers were built before: every anonymous counter starts somewhere else,
    # nested layers consume numbers in between (gaps), and the file is in arbitrary order
    exported = K.export_structural(m, keras_counters={'conv2d': 7, 'conv_block': 2, 'dense': 1})
    keys = list(exported)
    rng = np.random.default_rng(0)
    rng.shuffle(keys)
    fresh = _models()[which].to('cpu').init_weights(99)
    mapping = K.load_keras_weights(fresh, {k: exported[k] for k in keys})
    assert set(mapping) == set(ref)
    for k, v in fresh.get_weights().items():
        np.testing.assert_array_equal(v, ref[k])


def test_names_follow_the_reference():
    """Spot checks of the structural names against dl4ds/models: explicit names are kept (sp_postups.py:148,205),
    anonymous layers are numbered per Keras type in construction order (:134,156,163,207-212), TimeDistributed adds
    'layer' (spt_postups.py:131), EncoderBlock holds its ConvBlock under 'conv' (blocks.py:609), DeconvolutionBlock
    names its transposed convolutions conv2dtranspose1 / 2 (blocks.py:508-516)."""
    ms = _models()
    e = K.export_structural(ms['cfg2'].to('cpu').init_weights(0))
    assert 'conv2d|kernel' in e and 'conv2d_1|kernel' in e                 # stem, backbone_last
    assert 'ResidualBlock2|conv1x1|kernel' in e and 'transition_block|conv|kernel' in e
    assert 'TransitionLast|conv|bias' in e and 'conv_block|att|conv1|kernel' in e and 'conv_block_1|conv2|kernel' in e
    assert 'SubpixelConvolution|conv2x|kernel' in e
    e = K.export_structural(ms['cfg3'].to('cpu').init_weights(0))
    assert 'Deconvolution|conv2dtranspose1|kernel' in e and 'Deconvolution|conv2dtranspose2|kernel' in e
    assert 'localized_conv_block|localconv|kernel' in e and 'DenseBlock1|norm1|moving_mean' in e
    e = K.export_structural(ms['cfg4'].to('cpu').init_weights(0))
    assert 'upsampling_rc|layer|conv|kernel' in e and 'RecurrentConvBlock2|convlstm1|recurrent_kernel' in e
    assert 'localized_conv_block|layer|transition|conv|kernel' in e
    e = K.export_structural(ms['cfg5'].to('cpu').init_weights(0))
    assert 'EncoderBlock1|conv|conv1|kernel' in e and 'ResizeConvolution1|conv|kernel' in e and 'Bottleneck|conv2|bias' in e


def test_errors():
    m = _models()['cfg2'].to('cpu').init_weights(0)
    e = K.export_structural(m)
    bad = dict(e)
    bad.pop('conv2d_1|kernel')
    with pytest.raises(KeyError):
        K.load_keras_weights(m, bad)
    bad = dict(e)
    bad['conv2d|kernel'] = np.zeros((3, 3, 2, 8), np.float32)
    with pytest.raises(ValueError):
        K.load_keras_weights(m, bad)
    bad = dict(e)
    bad['mystery|kernel'] = np.zeros(3, np.float32)
    with pytest.raises(KeyError):
        K.load_keras_weights(m, bad)
