"""Host-side helpers that sit on the boundary of the hot path: argument validation with the
reference's error conventions (dl4ds/utils.py:58-171), square cropping and cv2 resizing of numpy
grids (utils.py:251-401), the spatio-temporal <-> spatial sample reshapes (utils.py:20-45) and the
``Timing`` helper (utils.py:206-248).  Pure numpy / OpenCV -- nothing here runs on the GPU.
"""
import time
from datetime import timedelta

import cv2
import numpy as np

BACKBONE_BLOCKS = ['convnet', 'resnet', 'densenet', 'convnext', 'unet']
UPSAMPLING_METHODS = ['spc', 'rc', 'dc', 'pin']
POSTUPSAMPLING_METHODS = ['spc', 'rc', 'dc']
INTERPOLATION_METHODS = ['inter_area', 'nearest', 'bicubic', 'bilinear', 'lanczos']
LOSS_FUNCTIONS = ['mae', 'mse', 'dssim', 'dssim_mae', 'dssim_mse', 'dssim_mae_mse', 'msdssim',
                  'msdssim_mae', 'msdssim_mae_mse']
DROPOUT_VARIANTS = ['vanilla', 'gaussian', 'spatial', 'mcdrop', 'mcgaussiandrop', 'mcspatialdrop']

_CV2_MODES = {'nearest': cv2.INTER_NEAREST, 'bicubic': cv2.INTER_CUBIC, 'bilinear': cv2.INTER_LINEAR,
              'inter_area': cv2.INTER_AREA, 'lanczos': cv2.INTER_LANCZOS4}


# ---------------------------------------------------------------------------------------------
# argument checks (same conditions / exception types as the reference)
# ---------------------------------------------------------------------------------------------
def checkarg_upsampling(upsampling):
    if not isinstance(upsampling, str):
        raise TypeError('`upsampling` must be a string')
    if upsampling not in UPSAMPLING_METHODS:
        raise ValueError('`upsampling` not recognized. Must be one of the following: %s. Got %s'
                         % (UPSAMPLING_METHODS, upsampling))
    return upsampling


def checkarg_backbone(backbone):
    if not isinstance(backbone, str):
        raise TypeError('`backbone` must be a string')
    if backbone not in BACKBONE_BLOCKS:
        raise ValueError('`backbone` not recognized. Must be one of the following: %s. Got %s'
                         % (BACKBONE_BLOCKS, backbone))
    return backbone


def check_compatibility_upsbackb(backbone, upsampling, time_window):
    upsampling = checkarg_upsampling(upsampling)
    backbone = checkarg_backbone(backbone)
    if backbone == 'unet' and upsampling != 'pin':
        raise ValueError('`unet` backbone only works with `pin` pre-upsampling')
    if backbone in ('convnext', 'unet') and time_window is not None:
        raise ValueError('`unet` and `convnext` backbones only work with spatial samples '
                         '(`time_window` must be None)')
    return backbone, upsampling


def checkarg_dropout_variant(dropout_variant):
    if dropout_variant is None or dropout_variant == 'vanilla':
        return dropout_variant
    if isinstance(dropout_variant, str):
        if dropout_variant not in DROPOUT_VARIANTS:
            raise ValueError('`dropout_variant` must be None or one of %s, got %s'
                             % (DROPOUT_VARIANTS, dropout_variant))
        return dropout_variant


def checkarg_loss(loss):
    """Returns the loss NAME (the engine looks the kernels up by name: engine.LOSS_TERMS)."""
    if not isinstance(loss, str):
        raise TypeError('`loss` must be a string, one of %s' % (LOSS_FUNCTIONS,))
    if loss not in LOSS_FUNCTIONS:
        raise ValueError('`loss` must be one of %s, got %s' % (LOSS_FUNCTIONS, loss))
    return loss


def checkarray_ndim(array, ndim=3, add_axis_position=-1):
    if array.ndim < ndim:
        return np.expand_dims(array, axis=add_axis_position)
    return array


# ---------------------------------------------------------------------------------------------
# sample reshapes
# ---------------------------------------------------------------------------------------------
def spatial_to_spatiotemporal_samples(array, time_window):
    n, y, x, c = array.shape
    n_t = n - (time_window - 1)
    out = np.zeros((n_t, time_window, y, x, c))
    for i in range(n_t):
        out[i] = array[i:i + time_window]
    return out


def spatiotemporal_to_spatial_samples(array, time_window):
    if array.shape[1] != time_window:
        raise ValueError('`time_window` must be located in the second position '
                         '[n_samples, time_window, lat, lon, vars]')
    return np.concatenate([array[:, 0], array[-1, 1:]], axis=0)


# ---------------------------------------------------------------------------------------------
# crop / resize
# ---------------------------------------------------------------------------------------------
def crop_array(array, size, yx=None, position=False, exclude_borders=False, get_copy=False):
    """Square crop of a 2-5D grid ([y,x], [y,x,c], [t,y,x,c], [n,t,y,x,c]); random corner when
    ``yx`` is None (``np.random.randint(0, extent - size)``, upper bound exclusive as in the
    reference, utils.py:299-311)."""
    if array.ndim not in (2, 3, 4, 5):
        raise TypeError('Input array is not a 2D, 3D, or 4D ndarray')
    if not isinstance(size, int):
        raise TypeError('`Size` must be integer')
    ax = {2: 0, 3: 0, 4: 1, 5: 2}[array.ndim]
    ny, nx = array.shape[ax], array.shape[ax + 1]
    if size > ny or size > nx:
        raise ValueError('`Size` larger than the input image size')
    if yx is not None and isinstance(yx, tuple):
        y, x = yx
    elif exclude_borders:
        y = np.random.randint(1, ny - size - 1)
        x = np.random.randint(1, nx - size - 1)
    else:
        y = np.random.randint(0, ny - size)
        x = np.random.randint(0, nx - size)
    y1, x1 = int(y + size), int(x + size)
    if y < 0 or x < 0 or y1 > ny or x1 > nx:
        raise RuntimeError('Cropped image cannot be obtained with size=%s, y=%s, x=%s' % (size, y, x))
    index = [slice(None)] * array.ndim
    index[ax], index[ax + 1] = slice(y, y1), slice(x, x1)
    out = array[tuple(index)]
    if get_copy:
        out = out.copy()
    return (out, y, x) if position else out


def resize_array(array, newsize, interpolation='inter_area', squeezed=True, keep_dynamic_range=False):
    """cv2.resize of [y,x], [y,x,c] or [t,y,x,c] grids; ``newsize`` is (x, y).  4-D inputs come
    back as float64 (frames are written into an ``np.zeros`` buffer, utils.py:388-392)."""
    if interpolation not in INTERPOLATION_METHODS:
        raise ValueError('`interpolation` must be one of %s. Received %s' % (INTERPOLATION_METHODS, interpolation))
    if array.dtype in ['bool', 'int', 'int64']:
        array = array.astype('int')
        interpolation = 'nearest'
    mode = _CV2_MODES[interpolation]
    sx, sy = newsize
    if array.ndim in (2, 3):
        out = cv2.resize(array, (sx, sy), interpolation=mode)
        if out.ndim == 2 and array.ndim == 3:
            out = out[..., None]
    elif array.ndim == 4:
        nch = array.shape[-1]
        out = np.zeros((array.shape[0], sy, sx, nch))
        for i, frame in enumerate(array):
            r = cv2.resize(frame, (sx, sy), interpolation=mode)
            out[i] = r[..., None] if nch == 1 else r
    else:
        raise RuntimeError('Wrong dimensions, got %d' % array.ndim)
    if squeezed:
        out = np.squeeze(out)
    if keep_dynamic_range:
        out = np.clip(out, a_min=array.min(), a_max=array.max())
    return out


class Timing:
    """Wall-clock bookkeeping (utils.py:206-248)."""

    def __init__(self, verbose=True):
        self.verbose = verbose
        self.running_time = None
        self.checktimes = [time.time()]
        if self.verbose:
            print('-' * 80)
            print('Starting time: ' + time.strftime('%Y-%m-%d %H:%M:%S'))
            print('-' * 80)

    def checktime(self):
        self.checktimes.append(time.time())
        if self.verbose:
            print('Timing: ' + str(timedelta(seconds=self.checktimes[-1] - self.checktimes[-2])))

    def runtime(self):
        self.running_time = str(timedelta(seconds=time.time() - self.checktimes[0]))
        if self.verbose:
            print('-' * 80)
            print('Total running time: ' + self.running_time)
            print('-' * 80)
