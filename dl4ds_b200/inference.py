"""Predictor / predict -- dl4ds/inference.py:12-255: one batch of all samples built by
``create_batch_hr_lr`` (the reference does NO spatial tiling at inference), forward in chunks of
``batch_size`` on the GPU, time-collapse for spatio-temporal models, inverse scaling, optional save.
"""
import os

import numpy as np

from .dataloader import create_batch_hr_lr
from .utils import Timing, checkarray_ndim, resize_array, spatiotemporal_to_spatial_samples


class Predictor:
    """Predictor -- inference.py:12-106 (same arguments)."""

    def __init__(self, trainer, array, scale, array_in_hr=False, static_vars=None, predictors=None,
                 time_window=None, time_metadata=None, interpolation='inter_area', batch_size=64,
                 scaler=None, save_path=None, save_fname='y_hat.npy', return_lr=False, device='GPU'):
        self.trainer = trainer
        self.array = array
        self.scale = scale
        self.array_in_hr = array_in_hr
        self.static_vars = static_vars
        self.predictors = predictors
        self.time_window = time_window
        self.time_metadata = time_metadata
        self.interpolation = interpolation
        self.batch_size = batch_size
        self.scaler = scaler
        self.save_path = save_path
        self.save_fname = save_fname
        self.return_lr = return_lr
        self.device = device

    def run(self):
        return predict(
            trainer=self.trainer, array=self.array, scale=self.scale, array_in_hr=self.array_in_hr,
            static_vars=self.static_vars, predictors=self.predictors, time_window=self.time_window,
            time_metadata=self.time_metadata, interpolation=self.interpolation,
            batch_size=self.batch_size, scaler=self.scaler, save_path=self.save_path,
            save_fname=self.save_fname, return_lr=self.return_lr, device=self.device)


def predict(trainer, array, scale, array_in_hr=True, static_vars=None, predictors=None,
            time_window=None, time_metadata=None, interpolation='inter_area', batch_size=64,
            scaler=None, save_path=None, save_fname='y_hat.npy', return_lr=False, device='GPU'):
    """predict -- inference.py:109-255."""
    timing = Timing(verbose=False)
    if hasattr(trainer, 'model'):
        model = trainer.model
    elif hasattr(trainer, 'generator'):
        model = trainer.generator
    else:
        model = trainer
    upsampling = model.name.split('_')[-1]
    dim = len(model.input.shape)
    if dim == 5 and time_window is None:
        raise ValueError('`time_window` must be provided for spatiotemporal model')
    if device != 'GPU':
        raise ValueError("dl4ds_b200 runs inference on CUDA only (device='GPU')")
    time_metadata = None
    array = getattr(array, 'values', array)
    if static_vars is not None:
        static_vars = [getattr(v, 'values', v) for v in static_vars]
    n_samples = array.shape[0]
    if time_window is not None:
        n_samples -= time_window - 1
    if predictors is not None:
        predictors = np.concatenate([getattr(p, 'values', p) for p in predictors], axis=-1)
    if array_in_hr:
        array_hr, array_lr = array, None
    else:
        array = checkarray_ndim(array, 4, -1)
        hr_xy = (array.shape[2] * scale, array.shape[1] * scale)
        array_hr = resize_array(array, hr_xy, interpolation, squeezed=False)
        array_lr = array
    batch = create_batch_hr_lr(
        all_indices=np.arange(n_samples), index=0, array=array_hr, array_lr=array_lr,
        upsampling=upsampling, scale=scale, batch_size=n_samples, patch_size=None,
        time_window=time_window, static_vars=static_vars, predictors=predictors,
        interpolation=interpolation, time_metadata=time_metadata)
    if static_vars is not None:
        [batch_lr, batch_aux_hr], _ = batch
        inputs = [np.asarray(batch_lr, np.float32), np.asarray(batch_aux_hr, np.float32)]
    else:
        [batch_lr], _ = batch
        inputs = [np.asarray(batch_lr, np.float32)]
    out = model.predict(inputs, batch_size=batch_size, verbose=0)
    if out.ndim == 5 and time_window is not None:
        out = spatiotemporal_to_spatial_samples(out, time_window)
    if scaler is not None:
        out = scaler.inverse_transform(out)
    if save_path is not None and save_fname is not None:
        np.save(os.path.join(save_path, save_fname), out.astype('float32'))
    timing.runtime()
    if return_lr:
        return out, np.array(inputs[0])
    return out
