"""compute_metrics / compute_rmse / compute_correlation -- dl4ds/metrics.py:15-327 with the reductions on the GPU
(SURVEY.md section 8f row 4).

The reference walks the grid point by point on the host (joblib, one sklearn / scipy call per grid point or per
sample).  Here ``dl4ds_metrics_moments`` reduces the (N, H*W*C) pair once per direction into fp64 raw moments and
``dl4ds_ssim_index`` evaluates ``tf.image.ssim`` per sample; every metric of the reference is a closed form of
those.  Spearman's rank correlation needs a sort per sample: it stays on the host (``scipy.stats.spearmanr``, the
reference's own call, metrics.py:84-96).  Plots (ecubevis / seaborn, metrics.py:206-291) are outside the path: the
``.npy`` maps and ``metrics_summary.txt`` the reference writes next to them are written.

Reference behaviour kept on purpose: the per-grid-point "RMSE" map holds MEAN SQUARED errors (``rmse_per_px`` calls
``mean_squared_error`` with its default ``squared=True``, metrics.py:25-27) while the per-sample list holds root mean
squared errors (:29-30); grid points where the FIRST ground-truth field is exactly zero are skipped (NaN in the maps:
``np.where(y[0,:,:,0])``, :36,74); only channel 0 enters the per-grid-point maps.
"""
import os

import numpy as np

from .utils import Timing, checkarray_ndim


def _moments(y, y_hat, pair=True, point=True):
    """fp64 raw moments of two (N,H,W,C) arrays on the GPU: (pair (N,11) or None, point (H,W,C,7) or None)."""
    import torch
    from . import _lib
    if not torch.cuda.is_available():
        raise RuntimeError('dl4ds_b200.metrics needs a CUDA device (there is no CPU fallback)')
    dev = torch.device('cuda', torch.cuda.current_device())
    a = torch.as_tensor(np.ascontiguousarray(y, dtype=np.float32)).to(dev)
    b = torch.as_tensor(np.ascontiguousarray(y_hat, dtype=np.float32)).to(dev)
    n = a.shape[0]
    p = int(np.prod(a.shape[1:]))
    po = torch.empty((n, 11), dtype=torch.float64, device=dev) if pair else None
    pt = torch.empty((p, 7), dtype=torch.float64, device=dev) if point else None
    _lib.call('dl4ds_metrics_moments', a.data_ptr(), b.data_ptr(), n, p, po.data_ptr() if pair else None,
              pt.data_ptr() if point else None, torch.cuda.current_stream().cuda_stream)
    return (po.cpu().numpy() if pair else None,
            pt.cpu().numpy().reshape(tuple(a.shape[1:]) + (7,)) if point else None)


def _pearson_from_moments(m, n):
    """Pearson r from raw sums (sum y, sum yh, sum y^2, sum yh^2, sum y*yh = m[..., 2:7]) over n values."""
    sy, sh, syy, shh, syh = (m[..., k] for k in (2, 3, 4, 5, 6))
    cov = syh - sy * sh / n
    vy = syy - sy * sy / n
    vh = shh - sh * sh / n
    with np.errstate(invalid='ignore', divide='ignore'):
        return cov / np.sqrt(vy * vh)


def _prep(y, y_hat):
    y = np.asarray(getattr(y, 'values', y))
    y_hat = np.asarray(getattr(y_hat, 'values', y_hat))
    return checkarray_ndim(y, 4, -1), checkarray_ndim(y_hat, 4, -1)


def compute_rmse(y, y_hat, over='time', squared=False, n_jobs=40):
    """metrics.py:15-48.  over='time': (H,W) map of the MEAN SQUARED error of channel 0 over the samples (NaN where
    ``y[0,:,:,0] == 0``); over='space': list of per-sample (R)MSE over all grid values (``squared`` as in sklearn)."""
    y, y_hat = _prep(y, y_hat)
    n = y.shape[0]
    if over == 'time':
        _, pt = _moments(y, y_hat, pair=False)
        out = np.full(y.shape[1:3], np.nan, dtype=y.dtype if y.dtype.kind == 'f' else np.float64)
        valid = y[0, :, :, 0] != 0
        out[valid] = (pt[:, :, 0, 0] / n)[valid]
        return out
    elif over == 'space':
        pr, _ = _moments(y, y_hat, point=False)
        mse = pr[:, 0] / float(np.prod(y.shape[1:]))
        return list(mse if squared else np.sqrt(mse))


def compute_correlation(y, y_hat, over='time', mode='spearman', n_jobs=40):
    """metrics.py:51-97.  Pearson from the device moments; Spearman through scipy on the host (rank transform)."""
    y, y_hat = _prep(y, y_hat)
    n = y.shape[0]
    if mode not in ('spearman', 'pearson'):
        raise ValueError("`mode` must be 'spearman' or 'pearson'")
    if over == 'time':
        out = np.full(y.shape[1:3], np.nan, dtype=y.dtype if y.dtype.kind == 'f' else np.float64)
        valid = y[0, :, :, 0] != 0
        if mode == 'pearson':
            _, pt = _moments(y, y_hat, pair=False)
            out[valid] = _pearson_from_moments(pt[:, :, 0, :], n)[valid]
        else:
            from scipy.stats import spearmanr
            for yy, xx in zip(*np.where(valid)):
                out[yy, xx] = spearmanr(y[:, yy, xx, 0], y_hat[:, yy, xx, 0])[0]
        return out
    elif over == 'space':
        if mode == 'pearson':
            pr, _ = _moments(y, y_hat, point=False)
            return list(_pearson_from_moments(pr, float(np.prod(y.shape[1:]))))
        from scipy.stats import spearmanr
        return [spearmanr(y[i].ravel(), y_hat[i].ravel())[0] for i in range(n)]


def ssim_per_sample(y, y_hat, max_val):
    """tf.image.ssim(y, y_hat, max_val) (metrics.py:172-176) -> (N,) float32, on the device."""
    import torch
    from . import _lib
    dev = torch.device('cuda', torch.cuda.current_device())
    a = torch.as_tensor(np.ascontiguousarray(y, dtype=np.float32)).to(dev)
    b = torch.as_tensor(np.ascontiguousarray(y_hat, dtype=np.float32)).to(dev)
    n, h, w, c = a.shape
    out = torch.empty(n, dtype=torch.float32, device=dev)
    chunk = max(1, 65535 // c)                      # the SSIM kernels index (sample, channel) planes by blockIdx.z
    st = torch.cuda.current_stream().cuda_stream
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        nws = _lib.load().dl4ds_ssim_loss_workspace_floats(m, h, w, c, 1)
        if nws < 0:
            raise _lib.Dl4dsError('ssim_index: %s' % _lib.last_error())
        ws = torch.empty(nws, dtype=torch.float32, device=dev)
        _lib.call('dl4ds_ssim_index', a[s:s + m].data_ptr(), b[s:s + m].data_ptr(), m, h, w, c, float(max_val),
                  out[s:s + m].data_ptr(), ws.data_ptr(), st)
    return out.cpu().numpy()


def compute_metrics(y_test, y_test_hat, dpi=150, plot_size_px=1000, n_jobs=-1, scaler=None, mask=None,
                    save_path=None, verbose=True):
    """Temporal and spatial verification metrics of a downscaled array against the ground truth -- metrics.py:100-327
    (same arguments; ``dpi`` / ``plot_size_px`` / ``n_jobs`` are accepted and unused: no plots, no joblib).  Returns
    ``(temp_rmse_map, temp_pearson_corrmap, nmeanbias)`` like the reference; the scalar summary is kept on
    ``compute_metrics.last_summary`` and printed / appended to ``metrics_summary.txt``."""
    timing = Timing(verbose)
    y_test = np.asarray(getattr(y_test, 'values', y_test))
    y_test_hat = np.asarray(getattr(y_test_hat, 'values', y_test_hat))
    if y_test.ndim == 5:
        y_test = np.squeeze(y_test, -1)
        y_test_hat = np.squeeze(y_test_hat, -1)
    y_test = checkarray_ndim(y_test, 4, -1)
    y_test_hat = checkarray_ndim(y_test_hat, 4, -1)
    if scaler is not None and hasattr(scaler, 'inverse_transform'):
        y_test = scaler.inverse_transform(y_test)
        y_test_hat = scaler.inverse_transform(y_test_hat)
    mask_nan = None
    if mask is not None:
        mask = np.array(getattr(mask, 'values', mask), copy=True)
        if mask.ndim == 2:
            mask = np.expand_dims(mask, -1)
        y_test = y_test * mask
        y_test_hat = y_test_hat * mask
        mask_nan = mask.astype('float').copy()
        mask_nan[mask == 0] = np.nan
        mask = np.squeeze(mask)
    n = y_test.shape[0]
    per_sample = float(np.prod(y_test.shape[1:]))

    pr, pt = _moments(y_test, y_test_hat)
    # dynamic range, PSNR (tf.image.psnr: 20 log10(max_val) - 10 log10(mse)), SSIM, MAE  -- metrics.py:166-184
    drange = max(pr[:, 8].max(), pr[:, 10].max()) - min(pr[:, 7].min(), pr[:, 9].min())
    mse_pairs = pr[:, 0] / per_sample
    with np.errstate(divide='ignore'):
        psnr = 20.0 * np.log10(drange) - 10.0 * np.log10(mse_pairs)
    ssim = ssim_per_sample(y_test, y_test_hat, drange)
    maes_pairs = pr[:, 1] / per_sample
    # RMSE -- metrics.py:186-209
    valid = y_test[0, :, :, 0] != 0
    temp_rmse_map = np.full(y_test.shape[1:3], np.nan)
    temp_rmse_map[valid] = (pt[:, :, 0, 0] / n)[valid]
    spatial_rmse = np.sqrt(mse_pairs)
    mean_temp_rmse, std_temp_rmse = np.nanmean(temp_rmse_map), np.nanstd(temp_rmse_map)
    if mask is not None:
        temp_rmse_map[np.where(mask == 0)] = 0
    mean_y = pr[:, 2].sum() / (n * per_sample)
    norm_temp_rmse_map = temp_rmse_map / (mean_y * 100)
    norm_mean_temp_rmse, norm_std_temp_rmse = np.nanmean(norm_temp_rmse_map), np.nanstd(norm_temp_rmse_map)
    if mask is not None:
        norm_temp_rmse_map[np.where(mask == 0)] = 0
    # normalised mean bias -- metrics.py:225-239
    nmeanbias = (pt[..., 3] - pt[..., 2]) / n
    nmeanbias = nmeanbias / (mean_y * 100)
    if mask is not None:
        nmeanbias = nmeanbias * mask_nan
    mean_nmeanbias = np.nanmean(nmeanbias)
    if mask is not None:
        nmeanbias[np.where(mask == 0)] = 0
    # correlations -- metrics.py:241-266
    spatial_spearman_corr = compute_correlation(y_test, y_test_hat, over='space')
    spatial_pearson_corr = list(_pearson_from_moments(pr, per_sample))
    temp_pearson_corrmap = np.full(y_test.shape[1:3], np.nan)
    temp_pearson_corrmap[valid] = _pearson_from_moments(pt[:, :, 0, :], n)[valid]
    mean_temp_pearson_corr, std_temp_pearson_corr = np.nanmean(temp_pearson_corrmap), np.nanstd(temp_pearson_corrmap)
    if mask is not None:
        temp_pearson_corrmap[np.where(mask == 0)] = 0

    if save_path is not None:
        sv = lambda name, a: np.save(os.path.join(save_path, name), a)
        sv('metrics_mse_pergridpair.npy', spatial_rmse)
        sv('metrics_pergridpoint_rmse_map.npy', temp_rmse_map)
        sv('metrics_pergridpoint_nrmse_map.npy', norm_temp_rmse_map)
        sv('metrics_nmeanbias_map.npy', nmeanbias)
        sv('metrics_spearcorr_pergridpair.npy', spatial_spearman_corr)
        sv('metrics_pearcorr_pergridpair.npy', spatial_pearson_corr)
        sv('metrics_pergridpoint_corrpears_map.npy', temp_pearson_corrmap)

    s = {
        'psnr': (np.mean(psnr), np.std(psnr)), 'ssim': (np.mean(ssim), np.std(ssim)),
        'mae': (np.mean(maes_pairs), np.std(maes_pairs)),
        'temp_rmse': (mean_temp_rmse, std_temp_rmse), 'temp_nrmse': (norm_mean_temp_rmse, norm_std_temp_rmse),
        'spatial_rmse': (np.mean(spatial_rmse), np.std(spatial_rmse)),
        'spatial_spearman': (np.mean(spatial_spearman_corr), np.std(spatial_spearman_corr)),
        'spatial_pearson': (np.mean(spatial_pearson_corr), np.std(spatial_pearson_corr)),
        'temp_pearson': (mean_temp_pearson_corr, std_temp_pearson_corr), 'nmeanbias': (mean_nmeanbias, None),
        'drange': drange, 'psnr_values': psnr, 'ssim_values': ssim, 'mae_values': maes_pairs,
    }
    compute_metrics.last_summary = s
    f = open(os.path.join(save_path, 'metrics_summary.txt'), 'a') if save_path is not None else None
    if f is not None or verbose:
        # the lines (and their labels, including the two Spearman rows) follow metrics.py:305-317
        print('Metrics on y_test and y_test_hat:\n', file=f)
        print('PSNR \tmu = %s \tsigma = %s' % s['psnr'], file=f)
        print('SSIM \tmu = %s \tsigma = %s' % s['ssim'], file=f)
        print('MAE \tmu = %s \tsigma = %s' % s['mae'], file=f)
        print('Per-grid-point RMSE \tmu = %s \tsigma = %s' % s['temp_rmse'], file=f)
        print('Per-grid-point nRMSE \tmu = %s \tsigma = %s' % s['temp_nrmse'], file=f)
        print('Per-grid-point Spearman correlation \tmu = %s \tsigma = %s' % s['spatial_spearman'], file=f)
        print('Per-grid-point Pearson correlation \tmu = %s \tsigma = %s' % s['temp_pearson'], file=f)
        print(file=f)
        print('Spatial MSE \tmu = %s \tsigma = %s' % s['spatial_rmse'], file=f)
        print('Spatial Spearman correlation \tmu = %s \tsigma = %s' % s['spatial_spearman'], file=f)
        print('Spatial Pearson correlation \tmu = %s \tsigma = %s' % s['spatial_pearson'], file=f)
    if f is not None:
        f.close()
    timing.runtime()
    return temp_rmse_map, temp_pearson_corrmap, nmeanbias


compute_metrics.last_summary = None
