"""GPU parity of the SSIM-family losses (dl4ds_b200/csrc/ssim.cu through the C ABI) against the oracle
(oracle/torch_ref.py, evaluated in fp64): value and gradient for every entry of LOSS_FUNCTIONS, the
SupervisedTrainer step with such a loss, and size-independent checks at the headline batch size."""
import numpy as np
import pytest
import torch

from dl4ds_b200 import SupervisedTrainer, losses
from dl4ds_b200._lib import Dl4dsError
from dl4ds_b200.engine import LOSS_TERMS
from oracle import torch_ref as R

pytestmark = pytest.mark.gpu

# tolerances: loss value 2e-5 absolute (fp32 sums of ~1e6 window positions); gradient 2e-4 of its largest
# regular element; the arg-max / arg-min elements of y_pred carry the sums over the whole batch that flow through
# the dynamic range and the shift (fp32 atomics, heavy cancellation): 2 % of their own size.
LOSS_TOL, GRAD_TOL, SPECIAL_TOL = 2e-5, 2e-4, 2e-2


def _pair(rng, shape, kind):
    yt = rng.standard_normal(shape)
    yp = yt + 0.3 * rng.standard_normal(shape)
    if kind == 'true_range':
        yt = yt * 3
    elif kind == 'positive':
        yt, yp = np.abs(yt), np.abs(yp) + 0.1
    return yt.astype(np.float32), yp.astype(np.float32)


def _oracle64(name, yt, yp):
    a = torch.tensor(yt, dtype=torch.float64)
    p = torch.tensor(yp, dtype=torch.float64, requires_grad=True)
    loss = R.LOSSES[name](a, p)
    loss.backward()
    return float(loss.detach()), p.grad.numpy()


def _check(name, yt, yp):
    ref, gref = _oracle64(name, yt, yp)
    out, dy = losses.value_and_grad(name, yt, yp)
    torch.cuda.synchronize()
    val, g = float(out.item()), dy.cpu().numpy().astype(np.float64).reshape(gref.shape)
    assert abs(val - ref) <= LOSS_TOL, (name, val, ref)
    special = [int(np.argmax(yp)), int(np.argmin(yp))]
    mask = np.ones(g.size, bool)
    mask[special] = False
    gf, rf = g.reshape(-1), gref.reshape(-1)
    scale = np.abs(rf[mask]).max()
    err = np.abs(gf[mask] - rf[mask]).max()
    assert err <= GRAD_TOL * scale, (name, err, scale)
    for i in special:
        assert abs(gf[i] - rf[i]) <= SPECIAL_TOL * abs(rf[i]) + GRAD_TOL * scale, (name, i, gf[i], rf[i])
    return val


@pytest.mark.parametrize('kind', ['plain', 'true_range', 'positive'])
@pytest.mark.parametrize('name', sorted(LOSS_TERMS))
def test_loss_and_gradient_match_oracle(cuda, name, kind):
    rng = np.random.default_rng(sum(map(ord, name + kind)))
    shapes = [(3, 96, 96, 2), (2, 128, 160, 1), (2, 90, 101, 1)]     # the last: odd sizes at scales 1-3 of MS-SSIM
    if 'ms' not in name:
        shapes.append((2, 24, 37, 1))            # ragged tiles, odd sizes (single scale only)
    for shape in shapes:
        yt, yp = _pair(rng, shape, kind)
        _check(name, yt, yp)


def test_spatiotemporal_samples_fold_onto_the_batch(cuda):
    rng = np.random.default_rng(5)
    yt, yp = _pair(rng, (2, 3, 32, 32, 1), 'plain')
    ref = float(R.dssim_mae(torch.tensor(yt).reshape(6, 32, 32, 1), torch.tensor(yp).reshape(6, 32, 32, 1)))
    assert abs(losses.dssim_mae(yt, yp) - ref) <= LOSS_TOL


def test_module_functions_and_errors(cuda):
    rng = np.random.default_rng(6)
    yt, yp = _pair(rng, (2, 96, 96, 1), 'plain')
    for name in LOSS_TERMS:
        ref = float(R.LOSSES[name](torch.tensor(yt, dtype=torch.float64), torch.tensor(yp, dtype=torch.float64)))
        assert abs(getattr(losses, name)(yt, yp) - ref) <= LOSS_TOL, name
    with pytest.raises(Dl4dsError):
        losses.dssim(yt[:, :10], yp[:, :10])          # smaller than the 11x11 window (tf.image.ssim asserts too)
    with pytest.raises(Dl4dsError):
        losses.msdssim(yt[:, :80, :80], yp[:, :80, :80])   # 80 -> 40 -> 20 -> 10 < 11: tf.image.ssim asserts too
    with pytest.raises(ValueError):
        losses.value_and_grad('ssim', yt, yp)


def test_fullsize_properties(cuda):
    """Headline batch (64 x 128 x 128 x 1): SSIM(x, x) = 1, symmetry, and a directional derivative against the
    returned gradient (size-independent checks; the oracle is not run at this size)."""
    rng = np.random.default_rng(7)
    yt, yp = _pair(rng, (64, 128, 128, 1), 'plain')
    for name in ('dssim', 'msdssim'):
        assert abs(getattr(losses, name)(yt, yt)) <= 1e-5
        assert abs(getattr(losses, name)(yt, yp) - getattr(losses, name)(yp, yt)) <= 1e-6
    d = rng.standard_normal(yp.shape).astype(np.float32)
    for name in ('dssim', 'msdssim', 'dssim_mse'):      # smooth terms only: MAE has kinks within +-eps
        _, g = losses.value_and_grad(name, yt, yp)
        slope = float((g.cpu().numpy().astype(np.float64) * d).sum())
        eps = 2e-2
        lp = float(losses.value_and_grad(name, yt, yp + eps * d, want_grad=False)[0].item())
        lm = float(losses.value_and_grad(name, yt, yp - eps * d, want_grad=False)[0].item())
        fd = (lp - lm) / (2 * eps)
        assert abs(fd - slope) <= 0.05 * abs(slope) + 1e-6, (name, fd, slope)


@pytest.mark.parametrize('loss,hw', [('dssim_mae', 64), ('msdssim_mae_mse', 96)])
def test_supervised_steps_with_ssim_losses(cuda, loss, hw):
    """Three optimizer steps through SupervisedTrainer(loss=...) == the oracle's supervised_step."""
    np.random.seed(0)
    hr = np.random.default_rng(3).standard_normal((16, hw, hw, 1)).astype(np.float32)
    tr = SupervisedTrainer('resnet', 'spc', hr, hr[:8], hr[:8], scale=4, batch_size=8, epochs=1, loss=loss,
                           learning_rate=(1e-3, 1e-4), lr_decay_after=100, verbose=False, math='fp32', seed=7,
                           n_blocks=2)
    tr.setup_datagen()
    tr.setup_model()
    w = {k: torch.from_numpy(v.copy()) for k, v in tr.model.get_weights().items()}
    opt = R.TFAdam(list(w), lr=R.piecewise_constant(100, 1e-3, 1e-4))
    fwd = lambda p, xs: R.net_postupsampling(p, xs, 'resnet', 'spc', 4, n_blocks=2)
    for i in range(3):
        (lr,), (y,) = tr.ds_train[i % len(tr.ds_train)]
        got = tr.train_on_batch([lr], y)
        ref, _ = R.supervised_step(fwd, w, opt, [torch.from_numpy(lr)], torch.from_numpy(y), loss=loss)
        assert abs(got - ref) <= 2e-4 * max(1.0, abs(ref)), (i, got, ref)
    from tests.util import assert_adam_weights_close
    assert_adam_weights_close(tr.model.get_weights(), {k: v.numpy() for k, v in w.items()}, lr=1e-3, steps=3,
                              tight=5e-5, frac=5e-3)
