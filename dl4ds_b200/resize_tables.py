"""Host-side tap tables for ``dl4ds_resample_taps``: a separable linear resampling is two (n_out, K) tables of source
indices and fp32 weights.

* :func:`tf_resize_matrix` -- the 1-D operator of ``tf.image.resize`` (what Keras ``Resizing(interpolation=...)`` runs,
  blocks.py:457-491) for the methods beyond bilinear / nearest / bicubic: ``area`` (ResizeArea op: box average of the
  source footprint) and ``lanczos3``, ``lanczos5``, ``gaussian``, ``mitchellcubic`` (ScaleAndTranslate op with
  ``antialias=False``: kernel support of fixed radius around the half-pixel sample position, weights renormalised
  over the taps that fall inside the image).  TensorFlow is third-party code that is not installable here; the
  formulas restate tensorflow/core/kernels/image/{scale_and_translate_op.cc, sampling_kernels.h, resize_area_op.cc}
  (TF 2.6-2.15), in float32 like the op.
* :func:`matrix_to_taps` -- dense operator -> padded (index, weight) tables.
"""
import math

import numpy as np

TAP_METHODS = ('area', 'lanczos3', 'lanczos5', 'gaussian', 'mitchellcubic')
_F = np.float32


def _kernel(method):
    """(radius, f(|x|) -> weight) of ScaleAndTranslate's sampling kernels (sampling_kernels.h)."""
    if method in ('lanczos3', 'lanczos5'):
        r = 3.0 if method == 'lanczos3' else 5.0

        def f(x):
            if x > r:
                return 0.0
            if x <= 1e-3:                    # the limit of sin(x) / x
                return 1.0
            pi = 3.14159265359
            return r * math.sin(pi * x) * math.sin(pi * x / r) / (pi * pi * x * x)
        return r, f
    if method == 'gaussian':                # GaussianKernelFunc(radius 1.5): sigma = radius / 3
        r, sigma = 1.5, 0.5
        return r, lambda x: 0.0 if x >= r else math.exp(-x * x / (2.0 * sigma * sigma))
    if method == 'mitchellcubic':           # Mitchell-Netravali, B = C = 1/3
        def f(x):
            if x >= 2.0:
                return 0.0
            if x >= 1.0:
                return (((-7.0 / 18.0) * x + 2.0) * x - 10.0 / 3.0) * x + 16.0 / 9.0
            return (((7.0 / 6.0) * x - 2.0) * x) * x + 8.0 / 9.0
        return 2.0, f
    raise ValueError(method)


def tf_resize_matrix(n_in, n_out, method):
    """(n_out, n_in) float32 operator of ``tf.image.resize(..., method, antialias=False)`` along one axis."""
    m = np.zeros((n_out, n_in), _F)
    if method == 'area':                    # ResizeArea: scale = in / out, cells [x*scale, (x+1)*scale)
        scale = _F(n_in) / _F(n_out)
        for x in range(n_out):
            in_x, in_x1 = _F(x) * scale, _F(x + 1) * scale
            for i in range(int(math.floor(in_x)), int(math.ceil(in_x1))):
                if i < in_x:
                    w = scale if i + 1 > in_x1 else _F(i + 1) - in_x
                else:
                    w = in_x1 - _F(i) if i + 1 > in_x1 else _F(1.0)
                m[x, min(max(i, 0), n_in - 1)] += _F(w) / scale
        return m
    radius, kern = _kernel(method)
    inv_scale = _F(1.0) / (_F(n_out) / _F(n_in))          # ComputeSpansCore: scale = out / in, antialias off
    for x in range(n_out):
        sample = (_F(x) + _F(0.5)) * inv_scale
        if sample < 0 or sample > n_in:
            continue
        lo = min(max(int(math.ceil(sample - radius - 0.5)), 0), n_in - 1)
        hi = min(max(int(math.floor(sample + radius - 0.5)), 0), n_in - 1) + 1
        w = np.array([kern(abs(float(_F(s) + _F(0.5) - sample))) for s in range(lo, hi)], _F)
        tot = _F(w.sum(dtype=_F))
        if abs(tot) >= 1000.0 * np.finfo(_F).tiny:
            m[x, lo:hi] = w * (_F(1.0) / tot)
    return m


def matrix_to_taps(r):
    """Dense (n_out, n_in) operator -> (indices int32 (n_out, K), weights float32 (n_out, K)), zero padded."""
    n_out = r.shape[0]
    K = int(max(1, (r != 0).sum(axis=1).max()))
    idx = np.zeros((n_out, K), np.int32)
    wts = np.zeros((n_out, K), np.float32)
    for o in range(n_out):
        nz = np.nonzero(r[o])[0]
        idx[o, :len(nz)] = nz
        wts[o, :len(nz)] = r[o, nz]
    return idx, wts
