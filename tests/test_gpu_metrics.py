"""GPU compute_metrics / compute_rmse / compute_correlation (dl4ds_b200/metrics.py) against the CPU
restatement of dl4ds/metrics.py built on the reference's own sklearn / scipy calls (oracle/metrics_np.py), and the
absl command-line app end to end."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from oracle import metrics_np as M

pytestmark = pytest.mark.gpu


def _pair(n=7, h=40, w=36, c=1, seed=0):
    rng = np.random.default_rng(seed)
    y = rng.standard_normal((n, h, w, c)).astype(np.float32) + 2.0
    yh = (y + 0.3 * rng.standard_normal(y.shape)).astype(np.float32)
    y[0, 5:8, 3:6, 0] = 0.0                 # grid points the reference skips (np.where(y[0,:,:,0]))
    return y, yh


@pytest.mark.parametrize('c', [1, 2])
def test_rmse_and_correlation(cuda, c):
    from dl4ds_b200 import compute_correlation, compute_rmse
    y, yh = _pair(c=c)
    np.testing.assert_allclose(compute_rmse(y, yh, over='time'), M.compute_rmse(y, yh, over='time'), rtol=2e-6)
    np.testing.assert_allclose(compute_rmse(y, yh, over='space'), M.compute_rmse(y, yh, over='space'), rtol=2e-6)
    np.testing.assert_allclose(compute_rmse(y, yh, over='space', squared=True),
                               M.compute_rmse(y, yh, over='space', squared=True), rtol=2e-6)
    for mode in ('pearson', 'spearman'):
        for over in ('time', 'space'):
            a, b = compute_correlation(y, yh, over=over, mode=mode), M.compute_correlation(y, yh, over=over, mode=mode)
            np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6, equal_nan=True)


@pytest.mark.parametrize('masked', [False, True])
def test_compute_metrics(cuda, tmp_path, masked):
    from dl4ds_b200 import compute_metrics
    y, yh = _pair(n=9)
    mask = None
    if masked:
        mask = np.ones(y.shape[1:3])
        mask[:4, :] = 0
    rmse_map, corr_map, bias = compute_metrics(y, yh, mask=mask, save_path=str(tmp_path), verbose=False)
    ref = M.compute_metrics(y, yh, mask=mask)
    s = compute_metrics.last_summary
    np.testing.assert_allclose(rmse_map, ref['temp_rmse_map'], rtol=2e-6, equal_nan=True)
    np.testing.assert_allclose(corr_map, ref['temp_pearson_corrmap'], rtol=1e-5, atol=1e-6, equal_nan=True)
    np.testing.assert_allclose(bias, ref['nmeanbias'], rtol=1e-5, atol=1e-9, equal_nan=True)
    assert abs(s['drange'] - ref['drange']) <= 1e-6 * ref['drange']
    np.testing.assert_allclose(s['psnr_values'], ref['psnr'], rtol=1e-6)
    np.testing.assert_allclose(s['ssim_values'], ref['ssim'], rtol=2e-5)
    np.testing.assert_allclose(s['mae_values'], ref['mae'], rtol=1e-6)
    np.testing.assert_allclose(s['spatial_rmse'][0], np.mean(ref['spatial_rmse']), rtol=1e-6)
    np.testing.assert_allclose(s['spatial_pearson'][0], np.mean(ref['spatial_pearson']), rtol=1e-6)
    np.testing.assert_allclose(s['spatial_spearman'][0], np.mean(ref['spatial_spearman']), rtol=1e-6)
    np.testing.assert_allclose(s['temp_rmse'][0], ref['mean_temp_rmse'], rtol=1e-5)
    np.testing.assert_allclose(s['temp_nrmse'][0], ref['norm_mean_temp_rmse'], rtol=1e-5)
    np.testing.assert_allclose(s['temp_pearson'][0], ref['mean_temp_pearson'], rtol=1e-5)
    np.testing.assert_allclose(s['nmeanbias'][0], ref['mean_nmeanbias'], rtol=1e-4, atol=1e-9)
    for f in ('metrics_pergridpoint_rmse_map.npy', 'metrics_nmeanbias_map.npy', 'metrics_summary.txt',
              'metrics_pearcorr_pergridpair.npy', 'metrics_spearcorr_pergridpair.npy'):
        assert os.path.exists(tmp_path / f)


def test_cli_app_train_test_metrics(cuda, tmp_path):
    """python -m dl4ds_b200.app --flagfile=... (app.py:1-304): train, predict and verify in one run."""
    data = tmp_path / 'data_module.py'
    data.write_text(textwrap.dedent('''
        import numpy as np
        rng = np.random.default_rng(0)
        _f = lambda n: rng.standard_normal((n, 32, 32, 1)).astype(np.float32)
        data_train, data_val, data_test = _f(24), _f(8), _f(8)
        data_train_lr = data_val_lr = data_test_lr = None
        predictors_train = predictors_val = predictors_test = None
        static_vars = None
        inference_data = data_test
        inference_predictors = None
        inference_scaler = None
        gt_holdout_dataset = data_test
        gt_mask = None
    '''))
    out = tmp_path / 'results'
    cfg = tmp_path / 'params.cfg'
    cfg.write_text('\n'.join([
        '--data_module=%s' % data, '--backbone=resnet', '--upsampling=spc', '--scale=4', '--n_filters=4', '--n_blocks=2',
        '--dropout_rate=0', '--epochs=2', '--batch_size=4', '--save_path=%s/' % out, '--noshow_plot', '--noverbose',
        '--inference_array_in_hr', '--learning_rate=1e-3', '--learning_rate=1e-4']))
    env = dict(os.environ, PYTHONPATH=os.getcwd())
    r = subprocess.run([sys.executable, '-m', 'dl4ds_b200.app', '--flagfile=%s' % cfg], capture_output=True, text=True,
                       env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert (out / 'metrics_summary.txt').exists() and (out / 'y_hat.npy').exists()
    assert np.load(out / 'y_hat.npy').shape == (8, 32, 32, 1)
    assert 'PSNR' in (out / 'metrics_summary.txt').read_text()
