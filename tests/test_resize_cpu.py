"""CPU tests of the nearest / bicubic resampling restated for `rc_interpolation` (blocks.py:457-491 -> Keras Resizing
-> tf.image.resize(method, antialias=False)): the oracle's per-axis matrices against hand values, and against an
independent implementation of the same definition -- Pillow's float-image resize (Keys cubic a = -0.5, taps clipped
to the image and renormalised; nearest with half-pixel centres) -- at the integer upsampling factors the path uses."""
import numpy as np
import pytest
import torch
from PIL import Image

from dl4ds_b200 import nets
from oracle import torch_ref as R


def test_bicubic_weights_by_hand():
    m = R._resize_matrix(8, 16, 'bicubic')
    assert np.allclose(m.sum(axis=1), 1.0, atol=1e-12)
    # x2: interior outputs alternate between the offsets 0.75 and 0.25 of the Keys kernel (A = -0.5)
    w25 = np.array([-0.0703125, 0.8671875, 0.2265625, -0.0234375])
    assert np.allclose(m[6, 1:5], w25[::-1], atol=1e-7)      # out 6 -> 2.75: taps 1..4 at offset 0.75
    assert np.allclose(m[7, 2:6], w25, atol=1e-7)            # out 7 -> 3.25: taps 2..5 at offset 0.25
    # first output -> -0.25: taps -2, -1 leave the image (weight 0), taps 0, 1 are renormalised
    o0 = np.array([0.8671875, -0.0703125])
    assert np.allclose(m[0, :2], o0 / o0.sum(), atol=1e-7) and np.all(m[0, 2:] == 0)


def test_nearest_is_replication_at_integer_factors():
    m = R._resize_matrix(5, 20, 'nearest')
    assert np.array_equal(m, np.repeat(np.eye(5), 4, axis=0))


@pytest.mark.parametrize('method,pil', [('bicubic', Image.BICUBIC), ('nearest', Image.NEAREST), ('lanczos3', Image.LANCZOS)])
@pytest.mark.parametrize('scale', [2, 4])
def test_against_pillow(method, pil, scale):
    rng = np.random.default_rng(scale)
    x = rng.standard_normal((9, 13)).astype(np.float32)
    ref = np.asarray(Image.fromarray(x, mode='F').resize((13 * scale, 9 * scale), resample=pil), np.float64)
    got = R.resize(torch.tensor(x, dtype=torch.float64)[None, None], 9 * scale, 13 * scale, method)[0, 0].numpy()
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()


def test_builders_accept_the_methods():
    for method in ('nearest', 'bicubic'):
        m = nets.net_postupsampling('resnet', 'rc', 2, 1, 0, (8, 8), n_blocks=1, rc_interpolation=method)
        p = R.Params()
        y = R.net_postupsampling(p, [torch.zeros(1, 8, 8, 1)], 'resnet', 'rc', 2, n_blocks=1, rc_interpolation=method)
        assert dict(m.spec) == dict(p.spec) and tuple(y.shape) == (1, 16, 16, 1)
    with pytest.raises(NotImplementedError):
        nets.net_postupsampling('resnet', 'rc', 2, 1, 0, (8, 8), rc_interpolation='spline')


def test_scale_and_translate_methods():
    """area / lanczos3 / lanczos5 / gaussian / mitchellcubic (tf.image.resize without antialiasing): the product's
    tables (dl4ds_b200/resize_tables.py, float32 like the TF op) against the oracle's float64 statement; rows sum to
    one; area = replication and OpenCV's INTER_AREA at integer upsampling factors; the reference Mitchell-Netravali
    values (B = C = 1/3: k(0) = 8/9, k(1) = 1/18)."""
    import cv2
    from dl4ds_b200.resize_tables import TAP_METHODS, tf_resize_matrix
    for method in TAP_METHODS:
        for n_in, n_out in ((8, 16), (9, 36), (13, 65), (7, 10)):
            a, b = tf_resize_matrix(n_in, n_out, method), R._scale_and_translate_matrix(n_in, n_out, method)
            assert np.abs(a - b).max() <= 2e-6, (method, n_in, n_out)
            assert np.allclose(b.sum(axis=1), 1.0, atol=1e-12)
    assert np.array_equal(tf_resize_matrix(5, 20, 'area'), np.repeat(np.eye(5, dtype=np.float32), 4, axis=0))
    x = np.random.default_rng(1).standard_normal((6, 7)).astype(np.float32)
    got = R.resize(torch.tensor(x)[None, None], 18, 21, 'area')[0, 0].numpy()
    assert np.array_equal(got, cv2.resize(x, (21, 18), interpolation=cv2.INTER_AREA))
    m = R._scale_and_translate_matrix(8, 8, 'mitchellcubic')      # identity scale: samples sit on the source centres
    assert abs(m[4, 4] / m[4, 3] - (8.0 / 9.0) / (1.0 / 18.0)) < 1e-9


def test_bilinear_against_opencv_and_pillow():
    """The bilinear Resizing of ResizeConvolutionBlock (half-pixel centres, edge clamp, no antialiasing) against two
    independent implementations of the same definition at the upsampling factors the path uses."""
    import cv2
    rng = np.random.default_rng(3)
    x = rng.standard_normal((9, 13)).astype(np.float32)
    for scale in (2, 4):
        got = R.resize_bilinear(torch.tensor(x)[None, None], 9 * scale, 13 * scale)[0, 0].numpy()
        ref_cv = cv2.resize(x, (13 * scale, 9 * scale), interpolation=cv2.INTER_LINEAR)
        ref_pil = np.asarray(Image.fromarray(x, mode='F').resize((13 * scale, 9 * scale), resample=Image.BILINEAR))
        assert np.abs(got - ref_cv).max() <= 1e-5 * np.abs(ref_cv).max()
        assert np.abs(got - ref_pil).max() <= 1e-5 * np.abs(ref_pil).max()
