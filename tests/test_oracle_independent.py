"""Checks of the oracle's primitive ops against third-party implementations of the same definitions that happen to
be installed (scipy, einops) -- independent of both of this repo's own statements (oracle/torch_ref.py, ops_np.py)."""
import numpy as np
import pytest
import torch
from einops import rearrange
from scipy.signal import correlate2d

from oracle import torch_ref as R


@pytest.mark.parametrize('k', [1, 3, 5, 7])
def test_conv2d_same_against_scipy_correlate(k):
    """Keras Conv2D(padding='same', strides=1) = zero-padded cross-correlation summed over input channels."""
    rng = np.random.default_rng(k)
    x = rng.standard_normal((2, 3, 9, 11))                   # NCHW
    w = rng.standard_normal((k, k, 3, 4))                    # HWIO
    b = rng.standard_normal(4)
    y = R.conv2d(torch.tensor(x), torch.tensor(w), torch.tensor(b)).numpy()
    ref = np.zeros((2, 4, 9, 11))
    for n in range(2):
        for o in range(4):
            ref[n, o] = b[o] + sum(correlate2d(x[n, i], w[:, :, i, o], mode='same') for i in range(3))
    assert np.abs(y - ref).max() < 1e-10


@pytest.mark.parametrize('r', [2, 5])
def test_depth_to_space_against_einops(r):
    """tf.nn.depth_to_space on NHWC: channel index = (r1 * r + r2) * C_out + c."""
    rng = np.random.default_rng(r)
    x = rng.standard_normal((2, 4, 3, r * r * 3))            # NHWC
    ref = rearrange(x, 'b h w (r1 r2 c) -> b (h r1) (w r2) c', r1=r, r2=r)
    got = R._nhwc(R.depth_to_space(R._nchw(torch.tensor(x)), r)).numpy()
    assert np.array_equal(got, ref)


def test_block_mean_coarsening_against_opencv_inter_area():
    """HR -> LR at an integer factor: cv2.INTER_AREA (utils.py:376-384) is the s x s block mean the kernels compute."""
    import cv2
    rng = np.random.default_rng(2)
    x = rng.standard_normal((32, 48)).astype(np.float32)
    ref = cv2.resize(x, (12, 8), interpolation=cv2.INTER_AREA)
    got = x.reshape(8, 4, 12, 4).mean(axis=(1, 3))
    assert np.abs(got - ref).max() <= 1e-6
