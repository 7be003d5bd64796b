"""CPU tests of the SSIM-family losses (losses.py:23-151): the oracle's restatement of tf.image.ssim /
ssim_multiscale against an independent window-by-window numpy fp64 evaluation and against known answers, and the
closed-form backward that dl4ds_b200/csrc/ssim.cu implements against torch autograd of the oracle."""
import numpy as np
import pytest
import torch

from dl4ds_b200 import utils
from dl4ds_b200.engine import LOSS_TERMS, MSSSIM_POWER_FACTORS
from oracle import torch_ref as R
from tests import ssim_np


def _pair(rng, shape, kind):
    yt = rng.standard_normal(shape)
    yp = yt + 0.3 * rng.standard_normal(shape)
    if kind == 'true_range':        # max and min of the pair sit in y_true: no range gradient into y_pred
        yt = yt * 3
    elif kind == 'positive':        # nothing negative: no shift (losses.py:47-54 take the else branches)
        yt, yp = np.abs(yt), np.abs(yp) + 0.1
    return yt, yp


def test_gaussian_window_is_tf_fspecial():
    g = R._fspecial_gauss(11, 1.5, torch.float64).numpy()
    assert g.shape == (11, 11) and abs(g.sum() - 1) < 1e-12
    assert np.allclose(g, ssim_np.gauss2d(), atol=1e-15)
    assert np.allclose(g, np.outer(g.sum(1), g.sum(0)), atol=1e-15)      # separable: what the kernels exploit


def test_oracle_ssim_against_direct_windows():
    rng = np.random.default_rng(0)
    x = rng.random((14, 17))
    y = np.clip(x + 0.2 * rng.standard_normal(x.shape), 0, None)
    L = 1.7
    ref = ssim_np.ssim_direct(x, y, L)
    got = R.tf_image_ssim(torch.tensor(x)[None, :, :, None], torch.tensor(y)[None, :, :, None], L)
    assert abs(float(got) - ref) < 1e-12


def test_known_answers():
    rng = np.random.default_rng(1)
    x = torch.tensor(rng.standard_normal((2, 96, 96, 1)))
    assert abs(float(R.dssim(x, x))) < 1e-12                 # SSIM(x, x) = 1
    assert abs(float(R.msdssim(x, x))) < 1e-12
    y = torch.tensor(rng.standard_normal((2, 96, 96, 1)))
    d = float(R.dssim(x, y))
    assert 0.4 < d < 0.6                                     # independent noise: SSIM ~ 0
    assert abs(float(R.dssim_mae(x, y)) - (0.8 * d + 0.2 * float(R.mae(x, y)))) < 1e-12
    assert abs(float(R.dssim_mae_mse(x, y)) - (0.6 * d + 0.2 * float(R.mae(x, y)) + 0.2 * float(R.mse(x, y)))) < 1e-12
    # two channels: mean over channels of the per-channel SSIM
    x2, y2 = torch.cat([x, y], -1), torch.cat([y, y], -1)
    L = 3.0
    s = R.tf_image_ssim(x2, y2, L)
    s0 = R.tf_image_ssim(x2[..., :1], y2[..., :1], L)
    assert torch.allclose(s, (s0 + 1) / 2, atol=1e-12)


def test_loss_tables_cover_the_reference_list():
    assert set(LOSS_TERMS) == set(utils.LOSS_FUNCTIONS) == set(R.LOSSES)
    assert tuple(MSSSIM_POWER_FACTORS) == tuple(R._MSSSIM_POWER_FACTORS) == ssim_np.POWER_FACTORS
    for name, terms in LOSS_TERMS.items():
        assert abs(sum(w for _, w in terms) - 1.0) < 1e-12
        assert utils.checkarg_loss(name) == name
    with pytest.raises(ValueError):
        utils.checkarg_loss('ssim')
    with pytest.raises(TypeError):
        utils.checkarg_loss(None)


@pytest.mark.parametrize('multiscale,hw', [(False, 24), (False, 37), (True, 96), (True, 90), (True, 101)])
@pytest.mark.parametrize('kind', ['plain', 'true_range', 'positive'])
def test_closed_form_backward_equals_autograd(multiscale, hw, kind):
    """The kernel algorithm (three derivative maps, transposed gaussian filtering, pooling chain, arg-max / arg-min
    fix-ups for the dynamic range and the shift) == autograd of the oracle, in fp64."""
    rng = np.random.default_rng(hw)
    yt, yp = _pair(rng, (2, hw, hw), kind)
    a = torch.tensor(yt[..., None])
    p = torch.tensor(yp[..., None], requires_grad=True)
    loss = (R.msdssim if multiscale else R.dssim)(a, p)
    loss.backward()
    l2, g2 = ssim_np.loss_and_grad(yt, yp, multiscale)
    assert abs(float(loss.detach()) - l2) < 1e-12
    g = p.grad[..., 0].numpy()
    assert np.abs(g - g2).max() <= 1e-12 * max(1.0, np.abs(g).max())
