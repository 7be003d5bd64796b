"""Host data path (dl4ds_b200/dataloader.py, utils.py) against fixtures produced by the UNMODIFIED
reference's own create_batch_hr_lr / resize_array / crop_array / spatiotemporal_to_spatial_samples
(generated in the builder container by oracle/make_golden.py, which imports /root/reference under
TF stubs).  Bit-exact: this is float32 numpy/cv2 work with a fixed operation order."""
import os

import numpy as np
import pytest

from dl4ds_b200 import dataloader, utils
from oracle.make_golden import CASES, make_inputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.mark.parametrize('case', CASES)
def test_create_batch_matches_reference(case):
    kw = make_inputs(case)
    n = kw['array'].shape[0] - (kw.get('time_window') or 0)
    idx = np.arange(n)[::-1].copy()
    np.random.seed(7)
    ins, outs = dataloader.create_batch_hr_lr(idx, kw.pop('index'), **kw)
    with np.load(os.path.join(GOLDEN, 'datapath_%s.npz' % case)) as z:
        assert ins[0].dtype == np.float32 and outs[0].dtype == np.float32
        assert ins[0].shape == z['lr'].shape and np.array_equal(ins[0], z['lr'])
        assert outs[0].shape == z['hr'].shape and np.array_equal(outs[0], z['hr'])
        if 'aux' in z.files:
            assert len(ins) == 2 and np.array_equal(ins[1], z['aux'])
        else:
            assert len(ins) == 1


def test_utils_match_reference():
    rng = np.random.default_rng(99)
    a5 = rng.standard_normal((5, 3, 4, 4, 1)).astype(np.float32)
    img = rng.standard_normal((3, 8, 12, 2)).astype(np.float32)
    with np.load(os.path.join(GOLDEN, 'utils_misc.npz')) as z:
        assert np.array_equal(utils.spatiotemporal_to_spatial_samples(a5, 3), z['st2s'])
        for interp in utils.INTERPOLATION_METHODS:
            assert np.array_equal(utils.resize_array(img, (24, 16), interp, squeezed=False), z['resize_up_' + interp])
            assert np.array_equal(utils.resize_array(img, (6, 4), interp, squeezed=False), z['resize_dn_' + interp])
        assert np.array_equal(utils.crop_array(img, 4, yx=(2, 3)), z['crop'])


def test_inter_area_is_block_mean_in_cv2_order():
    """cv2.INTER_AREA at an integer factor == block mean accumulated as the CUDA coarsening kernel
    does (row-major window, groups of four: sum += ((v0+v1)+v2)+v3, then * 1/area)."""
    rng = np.random.default_rng(5)
    for s in (2, 4, 8):
        x = rng.standard_normal((16 * s // 2, 8 * s, 1)).astype(np.float32)
        ref = utils.resize_array(x, (x.shape[1] // s, x.shape[0] // s), 'inter_area', squeezed=False)
        h, w = x.shape[0] // s, x.shape[1] // s
        win = x[:, :, 0].reshape(h, s, w, s).transpose(0, 2, 1, 3).reshape(h, w, s * s)
        acc = np.zeros((h, w), np.float32)
        for k in range(0, s * s, 4):
            acc = acc + (((win[..., k] + win[..., k + 1]) + win[..., k + 2]) + win[..., k + 3])
        mine = acc * np.float32(1.0 / (s * s))
        if s == 2:      # cv2's 2x2 fast path pairs the taps differently: equal only to rounding
            assert np.abs(mine - ref[..., 0]).max() <= 1e-6
        else:
            assert np.array_equal(mine, ref[..., 0]), s


def test_datagenerator_len_and_shapes():
    rng = np.random.default_rng(0)
    hr = rng.standard_normal((10, 16, 16, 1)).astype(np.float32)
    g = dataloader.DataGenerator(hr, None, 'resnet', 'spc', 4, batch_size=4)
    assert len(g) == 2
    (lr,), (y,) = g[1]
    assert lr.shape == (4, 4, 4, 1) and y.shape == (4, 16, 16, 1)
    g = dataloader.DataGenerator(hr, None, 'resnet', 'rc', 4, batch_size=3, time_window=3)
    assert g.n == 7 and len(g) == 2
    with pytest.raises(ValueError):
        dataloader.DataGenerator(hr, None, 'resnet', 'spc', 4, batch_size=4, patch_size=6)


def test_argument_checks():
    with pytest.raises(TypeError):
        utils.checkarg_backbone(3)
    with pytest.raises(ValueError):
        utils.checkarg_upsampling('foo')
    with pytest.raises(ValueError):
        utils.check_compatibility_upsbackb('unet', 'spc', None)
    with pytest.raises(ValueError):
        utils.check_compatibility_upsbackb('unet', 'pin', 4)
    with pytest.raises(ValueError):
        utils.checkarg_loss('l1')
    with pytest.raises(ValueError):
        utils.crop_array(np.zeros((4, 4)), 5)
    with pytest.raises(ValueError):
        utils.resize_array(np.zeros((4, 4), np.float32), (2, 2), 'cubic')


def test_device_data_generator_domain():
    """What the device-resident data path takes over and what stays with the host generator."""
    from dl4ds_b200.dataloader import DeviceDataGenerator
    a = np.zeros((4, 30, 30, 1), np.float32)
    p_hr = [np.zeros((4, 30, 30, 1), np.float32)]
    ok = DeviceDataGenerator.supported
    assert not ok(a, a[:, :15, :15], 'spc', 2, 8, None, None, None, 'inter_area')      # explicit LR + post-upsampling patches
    assert ok(a, a[:, :15, :15], 'spc', 2, None, None, None, None, 'inter_area')      # explicit pairs (MOS)
    assert not ok(a, None, 'spc', 2, 16, None, None, p_hr, 'inter_area')        # post-upsampling patches + predictors (App. B #3)
    assert not ok(a, None, 'spc', 4, 18, None, None, None, 'inter_area')        # patch not divisible by the scale
    assert not ok(a, None, 'spc', 2, 30, None, None, None, 'inter_area')        # crop_array needs patch < grid
    assert not ok(a, None, 'pin', 2, 16, 3, None, None, 'inter_area')           # the reference's own (T,y) crop of squeezed 1-channel windows
    assert not ok(a, None, 'spc', 2, None, None, [np.zeros((31, 30))], None, 'inter_area')
    assert not ok(a, None, 'spc', 2, None, None, None, None, 'spline')
    assert ok(a, None, 'spc', 4, None, None, None, None, 'inter_area')          # 30 / 4: cv2's non-integer area path (tap tables)
    assert ok(a, None, 'pin', 2, None, None, None, None, 'inter_area')
    assert ok(a, None, 'spc', 2, None, None, None, None, 'bicubic')
    assert ok(a, None, 'rc', 2, None, 3, None, p_hr, 'inter_area')
    assert ok(a, None, 'spc', 2, None, None, None, None, 'inter_area')
