"""Loss functions of dl4ds/losses.py:5-151 evaluated on the B200 through the C ABI (``dl4ds_pixel_loss``,
``dl4ds_ssim_loss``).  Same names and ``(y_true, y_pred)`` argument order as the reference; inputs are NHWC host
arrays or CUDA fp32 tensors (5-D spatio-temporal samples are folded onto the batch axis: every loss here is a mean
over all leading axes); the value comes back as a Python float.  ``value_and_grad`` also returns d loss / d y_pred,
which is what the training steps seed their backward pass with (``engine.Ctx.loss``)."""
import numpy as np
import torch

from . import _lib
from .engine import LOSS_ACCUMULATE, LOSS_TERMS, MSSSIM_POWER_FACTORS, _PF_HOST, _stream


def _as_dev(a):
    if isinstance(a, torch.Tensor):
        t = a.to(device='cuda', dtype=torch.float32)
    else:
        t = torch.from_numpy(np.ascontiguousarray(getattr(a, 'values', a), dtype=np.float32)).cuda()
    if t.dim() == 5:
        t = t.reshape((-1,) + tuple(t.shape[2:]))
    if t.dim() != 4:
        raise ValueError('expected an NHWC (or NTHWC) array, got shape %s' % (tuple(t.shape),))
    return t.contiguous()


def value_and_grad(name, y_true, y_pred, want_grad=True, scale=1.0):
    """(loss value as a 1-element CUDA tensor, d loss / d y_pred or None) for any entry of LOSS_FUNCTIONS."""
    if name not in LOSS_TERMS:
        raise ValueError('unknown loss %r' % (name,))
    if not torch.cuda.is_available():
        raise RuntimeError('dl4ds_b200.losses needs a CUDA device (no CPU fallback)')
    yt, yp = _as_dev(y_true), _as_dev(y_pred)
    if yt.shape != yp.shape:
        raise ValueError('shape mismatch %s vs %s' % (tuple(yt.shape), tuple(yp.shape)))
    B, H, W, C = yp.shape
    out = torch.zeros(1, dtype=torch.float32, device=yp.device)
    dy = torch.empty_like(yp) if want_grad else None
    dyp = dy.data_ptr() if want_grad else None
    acc = 0
    for term, weight in LOSS_TERMS[name]:
        if term in ('mae', 'mse'):
            _lib.call('dl4ds_pixel_loss', yp.data_ptr(), yt.data_ptr(), out.data_ptr(), dyp, yp.numel(),
                      {'mae': 0, 'mse': 1}[term] | (LOSS_ACCUMULATE if acc else 0), float(scale * weight), _stream())
        else:
            n_scales = 1 if term == 'dssim' else len(MSSSIM_POWER_FACTORS)
            nws = _lib.load().dl4ds_ssim_loss_workspace_floats(B, H, W, C, n_scales)
            if nws < 0:
                raise _lib.Dl4dsError('ssim_loss: %s' % _lib.last_error())
            ws = torch.empty(nws, dtype=torch.float32, device=yp.device)
            _lib.call('dl4ds_ssim_loss', yp.data_ptr(), yt.data_ptr(), B, H, W, C, n_scales, _PF_HOST,
                      float(scale * weight), out.data_ptr(), dyp, acc, ws.data_ptr(), _stream())
        acc = 1
    return out, dy


def _value(name, y_true, y_pred):
    return float(value_and_grad(name, y_true, y_pred, want_grad=False)[0].item())


def mae(y_true, y_pred):
    """losses.py:5-11."""
    return _value('mae', y_true, y_pred)


def mse(y_true, y_pred):
    """losses.py:14-20."""
    return _value('mse', y_true, y_pred)


def dssim(y_true, y_pred):
    """losses.py:23-59."""
    return _value('dssim', y_true, y_pred)


def dssim_mae(y_true, y_pred):
    """losses.py:62-68."""
    return _value('dssim_mae', y_true, y_pred)


def dssim_mae_mse(y_true, y_pred):
    """losses.py:71-84."""
    return _value('dssim_mae_mse', y_true, y_pred)


def dssim_mse(y_true, y_pred):
    """losses.py:87-93."""
    return _value('dssim_mse', y_true, y_pred)


def msdssim(y_true, y_pred):
    """losses.py:96-131."""
    return _value('msdssim', y_true, y_pred)


def msdssim_mae(y_true, y_pred):
    """losses.py:134-140."""
    return _value('msdssim_mae', y_true, y_pred)


def msdssim_mae_mse(y_true, y_pred):
    """losses.py:143-151."""
    return _value('msdssim_mae_mse', y_true, y_pred)
