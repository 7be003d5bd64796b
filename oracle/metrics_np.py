"""TEST INFRASTRUCTURE -- CPU restatement of dl4ds/metrics.py (the numbers, not the plots), used only by tests/.

Follows the reference call by call: ``sklearn.metrics.mean_squared_error`` / ``scipy.stats.pearsonr`` /
``scipy.stats.spearmanr`` per grid point and per sample (metrics.py:15-97; both libraries are installed here, so these
ARE the reference's own third-party calls), ``tf.image.psnr`` as 20 log10(max_val) - 10 log10(mse) over (H,W,C) and
``tf.image.ssim`` through ``oracle.torch_ref.tf_image_ssim`` (TensorFlow itself is not installable: restated).
"""
import numpy as np
import torch
from scipy.stats import pearsonr, spearmanr
from sklearn.metrics import mean_squared_error

from . import torch_ref as R


def compute_rmse(y, y_hat, over='time', squared=False):
    """metrics.py:15-48."""
    if over == 'time':
        out = np.zeros_like(y[0, :, :, 0]) * np.nan
        for yy, xx in zip(*np.where(y[0, :, :, 0])):
            out[yy, xx] = mean_squared_error(y[:, yy, xx, 0], y_hat[:, yy, xx, 0])        # squared=True (default): MSE
        return out
    mse = [mean_squared_error(y[i].flatten(), y_hat[i].flatten()) for i in range(y.shape[0])]
    return mse if squared else [float(np.sqrt(m)) for m in mse]                           # squared=False: RMSE


def compute_correlation(y, y_hat, over='time', mode='spearman'):
    """metrics.py:51-97."""
    f = spearmanr if mode == 'spearman' else pearsonr
    if over == 'time':
        out = np.zeros_like(y[0, :, :, 0]) * np.nan
        for yy, xx in zip(*np.where(y[0, :, :, 0])):
            out[yy, xx] = f(y[:, yy, xx, 0], y_hat[:, yy, xx, 0])[0]
        return out
    return [f(y[i].ravel(), y_hat[i].ravel())[0] for i in range(y.shape[0])]


def compute_metrics(y_test, y_test_hat, mask=None):
    """metrics.py:133-266, the computed quantities as a dict (maps and per-sample lists included)."""
    if y_test.ndim == 5:
        y_test, y_test_hat = np.squeeze(y_test, -1), np.squeeze(y_test_hat, -1)
    if y_test.ndim < 4:
        y_test, y_test_hat = y_test[..., None], y_test_hat[..., None]
    mask_nan = None
    if mask is not None:
        mask = mask.copy()
        if mask.ndim == 2:
            mask = np.expand_dims(mask, -1)
        y_test, y_test_hat = y_test.copy(), y_test_hat.copy()
        for i in range(y_test.shape[0]):
            y_test[i] *= mask
            y_test_hat[i] *= mask
        mask_nan = mask.astype('float').copy()
        mask_nan[mask == 0] = np.nan
        mask = np.squeeze(mask)
    drange = max(y_test.max(), y_test_hat.max()) - min(y_test.min(), y_test_hat.min())
    d = y_test.astype(np.float64) - y_test_hat.astype(np.float64)
    mse = np.mean(d * d, axis=(1, 2, 3))
    psnr = 20.0 * np.log10(drange) - 10.0 * np.log10(mse)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32))          # NHWC
    ssim = R.tf_image_ssim(t(y_test), t(y_test_hat), float(drange)).numpy()
    maes_pairs = np.mean(np.mean(np.abs(d), axis=-1), axis=(1, 2))
    temp_rmse_map = compute_rmse(y_test, y_test_hat, over='time')
    spatial_rmse = compute_rmse(y_test, y_test_hat, over='space')
    out = {'drange': drange, 'psnr': psnr, 'ssim': ssim, 'mae': maes_pairs, 'spatial_rmse': np.asarray(spatial_rmse),
           'mean_temp_rmse': np.nanmean(temp_rmse_map), 'std_temp_rmse': np.nanstd(temp_rmse_map)}
    if mask is not None:
        temp_rmse_map[np.where(mask == 0)] = 0
    norm = temp_rmse_map / (np.mean(y_test) * 100)
    out['norm_mean_temp_rmse'] = np.nanmean(norm)
    nmeanbias = np.mean(y_test_hat - y_test, axis=0)
    nmeanbias = nmeanbias / (np.mean(y_test) * 100)
    if mask is not None:
        nmeanbias = nmeanbias * mask_nan
    out['mean_nmeanbias'] = np.nanmean(nmeanbias)
    if mask is not None:
        nmeanbias[np.where(mask == 0)] = 0
    out['spatial_spearman'] = np.asarray(compute_correlation(y_test, y_test_hat, over='space'))
    out['spatial_pearson'] = np.asarray(compute_correlation(y_test, y_test_hat, mode='pearson', over='space'))
    corrmap = compute_correlation(y_test, y_test_hat, mode='pearson')
    out['mean_temp_pearson'] = np.nanmean(corrmap)
    if mask is not None:
        corrmap[np.where(mask == 0)] = 0
    out.update(temp_rmse_map=temp_rmse_map, temp_pearson_corrmap=corrmap, nmeanbias=nmeanbias)
    return out
