"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- torch-CPU fp32 restatement of the
DL4DS network graphs, losses, optimizer and training steps.  PARITY UNPINNED (no reference
tests exist); semantics follow TF/Keras 2.x as listed in SURVEY.md App. A.

Every function cites the reference file:line (relative to /root/reference) it restates.
Layout conventions are the reference's: activations NHWC (NTHWC for spatio-temporal samples),
weights in Keras layout (Conv2D ``(kh,kw,Cin,Cout)``, Conv2DTranspose ``(kh,kw,Cout,Cin)``,
ConvLSTM2D ``(kh,kw,Cin,4F)`` / ``(kh,kw,F,4F)`` gate order i,f,c,o, LocallyConnected2D 1x1
``W[H,W,Cin,F]``, ``b[H,W,F]``, Dense ``(in,out)``).  Internally torch ops run NCHW.

Parameters are requested by name from a ``Params`` store while the forward runs; running a
forward in *spec mode* yields the ordered ``[(name, shape)]`` list (Keras creation order),
which the tests compare with the product's graph builder and with the notebook's summary table.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

POSTUPSAMPLING_METHODS = ['spc', 'rc', 'dc']


# ----------------------------------------------------------------------------------------------
# parameter store
# ----------------------------------------------------------------------------------------------
class Params:
    """Name -> tensor store.  ``Params()`` = spec mode (records shapes, hands out zeros);
    ``Params(weights)`` = run mode (hands out the given tensors, checks shapes)."""

    def __init__(self, weights=None, dtype=torch.float32, training=True):
        self.spec = OrderedDict()
        self.weights = weights
        self.dtype = dtype
        self.training = training        # Keras `training` flag: BatchNormalization batch vs moving statistics
        self.dropout_mask = None        # callable(call_index, NCHW shape, rate, variant) -> mask tensor (see dropout)
        self.n_dropout = 0

    def get(self, name, shape):
        shape = tuple(int(s) for s in shape)
        if name in self.spec:
            assert self.spec[name] == shape, (name, self.spec[name], shape)
        else:
            self.spec[name] = shape
        if self.weights is None:
            return torch.zeros(shape, dtype=self.dtype)
        w = self.weights[name]
        assert tuple(w.shape) == shape, (name, tuple(w.shape), shape)
        return w

    def n_params(self):
        return sum(int(np.prod(s)) for s in self.spec.values())


# ----------------------------------------------------------------------------------------------
# primitive ops (NCHW inside)
# ----------------------------------------------------------------------------------------------
def act(x, name):
    """tf.keras.layers.Activation(name) -- blocks.py:75."""
    if name is None or name == 'linear':
        return x
    if name == 'relu':
        return F.relu(x)
    if name == 'sigmoid':
        return torch.sigmoid(x)
    if name == 'tanh':
        return torch.tanh(x)
    if name == 'gelu':
        return F.gelu(x)  # exact erf form (Keras default approximate=False)
    raise NotImplementedError(name)


def same_pads(n, k, s):
    """TF 'SAME' padding: out = ceil(n/s); total = max((out-1)*s + k - n, 0); before = total//2."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def conv2d(x, w, b=None, stride=1, padding='same'):
    """Keras Conv2D (cross-correlation, HWIO kernel) -- blocks.py:49-61, discriminator.py:55-60."""
    kh, kw = w.shape[0], w.shape[1]
    wt = w.permute(3, 2, 0, 1)
    if padding == 'same':
        pt, pb = same_pads(x.shape[2], kh, stride)
        pl, pr = same_pads(x.shape[3], kw, stride)
        x = F.pad(x, (pl, pr, pt, pb))
    return F.conv2d(x, wt, b, stride=stride)


def conv2d_transpose_same(x, w, stride):
    """Keras Conv2DTranspose(f, k, strides=s, padding='same', use_bias=False) -- blocks.py:508-516.
    Defined as the input-gradient of the SAME conv: out[j] = sum_i in[i] * w[j - s*i + pad_before],
    out size = s*in.  Kernel layout (kh, kw, Cout, Cin)."""
    kh, kw = w.shape[0], w.shape[1]
    wt = w.permute(3, 2, 0, 1)  # torch conv_transpose2d weight: (Cin, Cout, kh, kw)
    full = F.conv_transpose2d(x, wt, stride=stride)
    ho, wo = x.shape[2] * stride, x.shape[3] * stride
    pt, _ = same_pads(ho, kh, stride)
    pl, _ = same_pads(wo, kw, stride)
    out = full[:, :, pt:pt + ho, pl:pl + wo]
    # k < s would leave the full result smaller than s*in; not used by DL4DS (k=9, s<=8 typical)
    assert out.shape[2] == ho and out.shape[3] == wo
    return out


def depth_to_space(x, r):
    """tf.nn.depth_to_space NHWC 'DCR' order -- blocks.py:427.
    out[n, h*r+i, w*r+j, c] = in[n, h, w, (i*r+j)*C + c]."""
    n, ch, h, w = x.shape
    c = ch // (r * r)
    x = x.view(n, r, r, c, h, w)           # (n, i, j, c, h, w)
    x = x.permute(0, 3, 4, 1, 5, 2)        # (n, c, h, i, w, j)
    return x.reshape(n, c, h * r, w * r)


def resize_bilinear(x, ho, wo):
    """keras Resizing(h, w, 'bilinear') = tf.image.resize half-pixel, no antialias
    -- blocks.py:489, discriminator.py:62."""
    return F.interpolate(x, size=(ho, wo), mode='bilinear', align_corners=False, antialias=False)


def _resize_matrix(n_in, n_out, method):
    """(n_out, n_in) resampling matrix of one axis for tf.image.resize(method, antialias=False) -- the TF kernels
    ResizeNearestNeighbor / ResizeBicubic with half_pixel_centers=True (tensorflow/core/kernels/image/
    resize_nearest_neighbor_op.cc, resize_bicubic_op.cc; third-party, restated from the published source):
    index arithmetic in float32 as TF does it."""
    m = np.zeros((n_out, n_in), np.float64)
    scale = np.float32(n_in) / np.float32(n_out)
    for o in range(n_out):
        if method == 'nearest':
            i = min(int(np.floor((np.float32(o) + np.float32(0.5)) * scale)), n_in - 1)
            m[o, i] = 1.0
            continue
        # bicubic: Keys kernel A=-0.5 through the 1024-entry coefficient table; taps outside the image get weight 0
        # and the rest is renormalised (GetWeightsAndIndices<HalfPixelScaler, true>)
        A = -0.5
        loc = (np.float32(o) + np.float32(0.5)) * scale - np.float32(0.5)
        fl = np.floor(loc)
        off = int(np.rint(np.float32(loc - fl) * np.float32(1024.0)))

        def near(t):
            x = np.float32(t) / np.float32(1024.0)
            return float(((A + 2) * x - (A + 3)) * x * x + 1)

        def far(t):
            x = np.float32(t) / np.float32(1024.0) + 1.0
            return float(((A * x - 5 * A) * x + 8 * A) * x - 4 * A)
        raw = [far(off), near(off), near(1024 - off), far(1024 - off)]
        w, idx = [], []
        for k in range(4):
            want = int(fl) - 1 + k
            got = min(max(want, 0), n_in - 1)
            idx.append(got)
            w.append(raw[k] if got == want else 0.0)
        tot = sum(w)
        if abs(tot) >= 1000.0 * np.finfo(np.float32).tiny:
            w = [v / tot for v in w]
        for i, v in zip(idx, w):
            m[o, i] += v
    return m


def _scale_and_translate_matrix(n_in, n_out, method):
    """tf.image.resize(method in lanczos3 / lanczos5 / gaussian / mitchellcubic, antialias=False) = the ScaleAndTranslate
    op (tensorflow/core/kernels/image/scale_and_translate_op.cc ComputeSpansCore + sampling_kernels.h; third-party,
    restated): per output sample the kernel is evaluated at the source centres inside [sample - radius, sample +
    radius] (clamped to the image) and the weights are normalised to sum 1.  'area' = the ResizeArea op
    (resize_area_op.cc): the box [x * s, (x + 1) * s) of source cells, partial cells weighted by their overlap."""
    m = np.zeros((n_out, n_in), np.float64)
    if method == 'area':
        s = n_in / n_out
        for o in range(n_out):
            a, b = o * s, (o + 1) * s
            i = int(np.floor(a))
            while i < np.ceil(b):
                overlap = min(i + 1, b) - max(i, a)
                m[o, min(max(i, 0), n_in - 1)] += overlap / s
                i += 1
        return m
    pi = 3.14159265359

    def lanczos(r):
        return lambda t: 0.0 if t > r else (1.0 if t <= 1e-3 else r * np.sin(pi * t) * np.sin(pi * t / r) / (pi * pi * t * t))
    kernels = {
        'lanczos3': (3.0, lanczos(3.0)),
        'lanczos5': (5.0, lanczos(5.0)),
        'gaussian': (1.5, lambda t: 0.0 if t >= 1.5 else np.exp(-t * t / (2.0 * 0.5 * 0.5))),
        'mitchellcubic': (2.0, lambda t: 0.0 if t >= 2.0 else (
            (((-7.0 / 18.0) * t + 2.0) * t - 10.0 / 3.0) * t + 16.0 / 9.0 if t >= 1.0 else
            (((7.0 / 6.0) * t - 2.0) * t) * t + 8.0 / 9.0)),
    }
    radius, k = kernels[method]
    for o in range(n_out):
        centre = (o + 0.5) * (n_in / n_out)
        if centre < 0 or centre > n_in:
            continue
        first = int(min(max(np.ceil(centre - radius - 0.5), 0), n_in - 1))
        last = int(min(max(np.floor(centre + radius - 0.5), 0), n_in - 1))
        w = np.array([k(abs(i + 0.5 - centre)) for i in range(first, last + 1)])
        if abs(w.sum()) >= 1000.0 * np.finfo(np.float32).tiny:
            m[o, first:last + 1] = w / w.sum()
    return m


def resize(x, ho, wo, method='bilinear'):
    """keras Resizing(ho, wo, interpolation=method) on NCHW ``x`` -- blocks.py:457-491."""
    if method == 'bilinear':
        return resize_bilinear(x, ho, wo)
    if method in ('area', 'lanczos3', 'lanczos5', 'gaussian', 'mitchellcubic'):
        rh = torch.as_tensor(_scale_and_translate_matrix(x.shape[2], ho, method), dtype=x.dtype)
        rw = torch.as_tensor(_scale_and_translate_matrix(x.shape[3], wo, method), dtype=x.dtype)
        return torch.einsum('oh,nchw,pw->ncop', rh, x, rw)
    if method not in ('nearest', 'bicubic'):
        raise NotImplementedError(method)
    rh = torch.as_tensor(_resize_matrix(x.shape[2], ho, method), dtype=x.dtype)
    rw = torch.as_tensor(_resize_matrix(x.shape[3], wo, method), dtype=x.dtype)
    return torch.einsum('oh,nchw,pw->ncop', rh, x, rw)


def maxpool2(x):
    """MaxPooling2D((2,2)) stride 2 'valid' -- blocks.py:613."""
    return F.max_pool2d(x, 2, 2)


def hard_sigmoid(x):
    """Keras 2.x hard_sigmoid = clip(0.2x+0.5, 0, 1) (ConvLSTM2D recurrent_activation default)."""
    return torch.clamp(0.2 * x + 0.5, 0.0, 1.0)


def local_conv1x1(x, w, b):
    """LocallyConnected2D(f, (1,1), implementation=3) -- blocks.py:322-328.
    out[n,h,w,o] = sum_c x[n,h,w,c] W[h,w,c,o] + b[h,w,o];  x NCHW here."""
    y = torch.einsum('nchw,hwco->nohw', x, w)
    return y + b.permute(2, 0, 1).unsqueeze(0)


# ----------------------------------------------------------------------------------------------
# blocks (dl4ds/models/blocks.py)
# ----------------------------------------------------------------------------------------------
def _conv(p, name, x, cout, k=3, bias=True, stride=1, padding='same'):
    cin = x.shape[1]
    w = p.get(name + '/kernel', (k, k, cin, cout))
    b = p.get(name + '/bias', (cout,)) if bias else None
    return conv2d(x, w, b, stride=stride, padding=padding)


def channel_attention(p, name, x, nf, r=4):
    """ChannelAttention2D -- blocks.py:537-593.  4-D: mean over (H,W).  For folded 5-D input the
    caller handles the (T,H) quirk (App. B #6) via ``channel_attention_5d``."""
    y = x.mean(dim=(2, 3), keepdim=True)
    y = F.relu(_conv(p, name + '/conv1', y, int(nf / r), k=1))
    y = torch.sigmoid(_conv(p, name + '/conv2', y, nf, k=1))
    return x * y


def channel_attention_5d(p, name, x5, nf, r=4):
    """ChannelAttention2D applied to (B,T,C,H,W) [= NTHWC in the reference]: tf.reduce_mean over
    axes [1,2] of NTHWC = (T,H) -- blocks.py:587 used at spt_postups.py:153-154 (App. B #6).
    Attention map has shape (B,1,1,W,C)."""
    b, t, c, h, w = x5.shape
    y = x5.mean(dim=(1, 3))                 # (B, C, W)
    y = y.unsqueeze(2)                      # (B, C, 1, W) as NCHW image of height 1
    y = F.relu(_conv(p, name + '/conv1', y, int(nf / r), k=1))
    y = torch.sigmoid(_conv(p, name + '/conv2', y, nf, k=1))   # (B, C, 1, W)
    return x5 * y.unsqueeze(1)              # broadcast over T and H


MC_DROPOUT = ('mcdrop', 'mcgaussiandrop', 'mcspatialdrop')


def dropout(p, x, rate, variant=None):
    """get_dropout_layer(rate, variant)(x) -- blocks.py:680-706: identity at rate 0; Dropout / GaussianDropout /
    SpatialDropout2D act in training mode only, their MC twins always (blocks.py:662-677).  All are ``x * mask``.
    TensorFlow's random stream cannot be restated, so the mask is SUPPLIED: ``p.dropout_mask(i, shape, rate,
    variant)`` returns the mask of the i-th dropout application (the parity tests hand over the masks the CUDA
    generator produced); without a supplier the op only counts itself (spec mode)."""
    if not rate or rate <= 0:
        return x
    if not (p.training or variant in MC_DROPOUT):
        return x
    p.n_dropout += 1
    if p.dropout_mask is None:
        return x
    return x * p.dropout_mask(p.n_dropout, tuple(x.shape), rate, variant)


def normalize(p, name, x, kind, eps=1e-3):
    """tf.keras.layers.BatchNormalization() / LayerNormalization() on NCHW ``x`` -- blocks.py:63-71.
    Keras defaults: axis=-1 (channels), epsilon=1e-3, BN momentum 0.99, gamma ones / beta zeros.
    BN, training: normalise with the batch mean / biased variance over (N,H,W); moving statistics updated in
    place (moving = moving*0.99 + batch*0.01, the moving variance with the unbiased batch variance, as the fused
    Keras path does).  BN, inference: moving statistics.  LN: per pixel over the channel axis."""
    c = x.shape[1]
    gamma = p.get(name + '/gamma', (c,)).view(1, c, 1, 1)
    beta = p.get(name + '/beta', (c,)).view(1, c, 1, 1)
    if kind == 'ln':
        mu = x.mean(dim=1, keepdim=True)
        var = ((x - mu) ** 2).mean(dim=1, keepdim=True)
        return (x - mu) / torch.sqrt(var + eps) * gamma + beta
    if kind != 'bn':
        raise ValueError('Normalization not supported, got %s' % kind)        # blocks.py:64-65
    mm = p.get(name + '/moving_mean', (c,))
    mv = p.get(name + '/moving_variance', (c,))
    if p.training:
        mu = x.mean(dim=(0, 2, 3))
        var = ((x - mu.view(1, c, 1, 1)) ** 2).mean(dim=(0, 2, 3))
        if p.weights is not None:
            m = x.numel() // c
            with torch.no_grad():
                mm.data.mul_(0.99).add_(0.01 * mu.detach())
                mv.data.mul_(0.99).add_(0.01 * var.detach() * (m / max(m - 1, 1)))
    else:
        mu, var = mm, mv
    return (x - mu.view(1, c, 1, 1)) / torch.sqrt(var.view(1, c, 1, 1) + eps) * gamma + beta


def conv_block(p, name, x, filters, activation='relu', attention=False, ks1=3, ks2=3, normalization=None,
               dropout_rate=0, dropout_variant=None):
    """ConvBlock.call -- blocks.py:87-103.  With a normalisation the convolutions carry no
    bias (blocks.py:37,44,52,58)."""
    nb = normalization is None
    y = dropout(p, x, dropout_rate, dropout_variant)
    y = _conv(p, name + '/conv1', y, filters, k=ks1, bias=nb)
    if not nb:
        y = normalize(p, name + '/norm1', y, normalization)
    y = act(y, activation)
    y = dropout(p, y, dropout_rate, dropout_variant)
    y = _conv(p, name + '/conv2', y, filters, k=ks2, bias=nb)
    if not nb:
        y = normalize(p, name + '/norm2', y, normalization)
    y = act(y, activation)
    if attention:
        y = channel_attention(p, name + '/att', y, filters)
    return y


def residual_block(p, name, x, filters, activation='relu', attention=False, use_1x1conv=False,
                   normalization=None, dropout_rate=0, dropout_variant=None):
    """ResidualBlock.call -- blocks.py:210-230."""
    nb = normalization is None
    y = dropout(p, x, dropout_rate, dropout_variant)
    y = _conv(p, name + '/conv1', y, filters, bias=nb)
    if not nb:
        y = normalize(p, name + '/norm1', y, normalization)
    y = act(y, activation)
    y = dropout(p, y, dropout_rate, dropout_variant)
    y = _conv(p, name + '/conv2', y, filters, bias=nb)
    if not nb:
        y = normalize(p, name + '/norm2', y, normalization)
    if attention:
        y = channel_attention(p, name + '/att', y, filters)
    if use_1x1conv:
        x = _conv(p, name + '/conv1x1', x, filters, k=1)
    return act(y + x, activation)


def dense_block(p, name, x, filters, activation='relu', attention=False, normalization=None, dropout_rate=0,
                dropout_variant=None):
    """DenseBlock.call -- blocks.py:262-277.  The pre-activation of X is discarded (:263-267,
    App. B #5): Y = conv3x3(act([norm2](conv1x1(X)))); out = concat([Y, X]).  DenseBlock re-creates conv1 / conv2
    WITH bias (:249-259); norm1 is applied to X and its result dropped, so its gamma / beta exist without a
    gradient and, for BN, its moving statistics still follow X."""
    if normalization is not None:
        normalize(p, name + '/norm1', x, normalization)
    y = _conv(p, name + '/conv1', x, 4 * filters, k=1)
    if normalization is not None:
        y = normalize(p, name + '/norm2', y, normalization)
    y = act(y, activation)
    y = dropout(p, y, dropout_rate, dropout_variant)     # dropout2 (:271-272); dropout1's output is dropped (:265-267)
    y = _conv(p, name + '/conv2', y, filters, k=3)
    if attention:
        y = channel_attention(p, name + '/att', y, filters)
    return torch.cat([y, x], dim=1)


def depthwise_conv2d(x, w, b):
    """tf.keras.layers.DepthwiseConv2D(k, padding='same', depth_multiplier=1): w (k,k,C,1), NCHW x."""
    k, c = w.shape[0], w.shape[2]
    pt, pb = same_pads(x.shape[2], k, 1)
    pl, pr = same_pads(x.shape[3], k, 1)
    wt = w.permute(2, 3, 0, 1)                          # (C,1,k,k)
    return F.conv2d(F.pad(x, (pl, pr, pt, pb)), wt, b, groups=c)


def convnext_block(p, name, x, filters, activation='gelu', normalization='ln', use_1x1conv=False, drop_path=0.,
                   layer_scale_init_value=0):
    """ConvNextBlock.call -- blocks.py:170-184.  ``layer_scale_init_value`` > 0: trainable per-channel ``gamma``
    multiplying the branch (:166-169,178-179); ``drop_path`` > 0: DropPath on the branch in training mode
    (:106-129,183: x / keep_prob * floor(keep_prob + U[0,1)) with one draw per sample -- the mask is supplied like
    the dropout masks, shape (N,1,1,1)).  LayerNormalization(epsilon=1e-6) / BatchNormalization() -- :161-164; any
    other value of ``normalization`` leaves the layer without ``self.norm`` and the reference's call() fails."""
    if normalization not in ('bn', 'ln'):
        raise ValueError('ConvNextBlock needs normalization bn or ln')
    c = x.shape[1]
    wd = p.get(name + '/dwconv/depthwise_kernel', (7, 7, c, 1))
    bd = p.get(name + '/dwconv/bias', (c,))
    y = depthwise_conv2d(x, wd, bd)
    y = normalize(p, name + '/norm', y, normalization, eps=1e-6 if normalization == 'ln' else 1e-3)
    w1 = p.get(name + '/pwconv1/kernel', (c, 4 * filters))
    b1 = p.get(name + '/pwconv1/bias', (4 * filters,))
    y = act(torch.einsum('nchw,cd->ndhw', y, w1) + b1.view(1, -1, 1, 1), activation)
    w2 = p.get(name + '/pwconv2/kernel', (4 * filters, filters))
    b2 = p.get(name + '/pwconv2/bias', (filters,))
    y = torch.einsum('nchw,cd->ndhw', y, w2) + b2.view(1, -1, 1, 1)
    if layer_scale_init_value > 0:
        y = p.get(name + '/gamma', (filters,)).view(1, -1, 1, 1) * y
    if use_1x1conv:
        x = _conv(p, name + '/conv1x1', x, filters, k=1)
    if drop_path and drop_path > 0:
        y = dropout(p, y, drop_path, 'droppath')
    return x + y


def transition_block(p, name, x, filters, activation='relu'):
    """TransitionBlock.call without BN: conv1x1 then activation -- blocks.py:306-308."""
    return act(_conv(p, name + '/conv', x, filters, k=1), activation)


def localized_conv_block(p, name, x, filters=2):
    """LocalizedConvBlock -- blocks.py:312-333."""
    y = transition_block(p, name + '/transition', x, filters)
    h, w = y.shape[2], y.shape[3]
    wk = p.get(name + '/localconv/kernel', (h, w, filters, filters))
    bk = p.get(name + '/localconv/bias', (h, w, filters))
    return local_conv1x1(y, wk, bk)


def subpixel_block(p, name, x, scale, n_filters):
    """SubpixelConvolutionBlock.call -- blocks.py:433-454 (shared conv2x across x2 stages)."""
    def up(x, factor):
        lname = {2: 'conv2x', 5: 'conv5x'}.get(factor, 'conv')
        x = _conv(p, name + '/' + lname, x, n_filters * factor * factor)
        return depth_to_space(x, factor)
    plan = {2: [2], 4: [2, 2], 8: [2, 2, 2], 10: [2, 5], 20: [2, 2, 5]}.get(scale, [scale])
    for f in plan:
        x = up(x, f)
    return x


def resize_conv_block(p, name, x, scale, n_filters, interpolation='bilinear'):
    """ResizeConvolutionBlock.call -- blocks.py:485-491."""
    ho, wo = int(x.shape[2] * scale), int(x.shape[3] * scale)
    return _conv(p, name + '/conv', resize(x, ho, wo, interpolation), n_filters)


def deconv_block(p, name, x, scale, n_filters, output_activation=None):
    """DeconvolutionBlock.call -- blocks.py:522-534, INCLUDING the if/if/else fall-through
    (App. B #4): scale 4 runs T1, T2 and then the stride-4 transpose (16x)."""
    def t(lname, x, stride, a):
        cin = x.shape[1]
        w = p.get(name + '/' + lname + '/kernel', (9, 9, n_filters, cin))
        return act(conv2d_transpose_same(x, w, stride), a)
    if scale == 4:
        x = t('deconv_1of2_scale_x2', x, 2, None)
        x = t('deconv_2of2_scale_x2', x, 2, output_activation)
    if scale == 8:
        x = t('deconv_1of2_scale_x2', x, 2, None)
        x = t('deconv_2of2_scale_x2', x, 2, output_activation)
        x = t('deconv_2of2_scale_x2', x, 2, output_activation)
    else:
        x = t('deconv_scale_x' + str(scale), x, scale, output_activation)
    return x


def convlstm2d(p, name, x5, filters, k):
    """ConvLSTM2D(filters, k, return_sequences=True, padding='same') -- blocks.py:350-355.
    x5: (B, T, C, H, W).  tanh / hard_sigmoid, gate order i,f,c,o, zero initial state."""
    b, t, cin, h, w = x5.shape
    wx = p.get(name + '/kernel', (k, k, cin, 4 * filters))
    wh = p.get(name + '/recurrent_kernel', (k, k, filters, 4 * filters))
    bias = p.get(name + '/bias', (4 * filters,))
    hs = torch.zeros(b, filters, h, w, dtype=x5.dtype)
    cs = torch.zeros(b, filters, h, w, dtype=x5.dtype)
    outs = []
    for ti in range(t):
        z = conv2d(x5[:, ti], wx, bias) + conv2d(hs, wh, None)
        zi, zf, zc, zo = torch.split(z, filters, dim=1)
        i = hard_sigmoid(zi)
        f = hard_sigmoid(zf)
        cs = f * cs + i * torch.tanh(zc)
        o = hard_sigmoid(zo)
        hs = o * torch.tanh(cs)
        outs.append(hs)
    return torch.stack(outs, dim=1)


def normalize_5d(p, name, x5, kind):
    """BatchNormalization / LayerNormalization on (B,T,C,H,W) [NTHWC in the reference], axis -1: batch statistics
    over (B,T,H,W), layer statistics over C -- the 4-D op on the tensor folded to (B*T,C,H,W)."""
    b, t = x5.shape[:2]
    return normalize(p, name, x5.reshape(b * t, *x5.shape[2:]), kind).reshape(x5.shape)


def recurrent_conv_block(p, name, x5, filters, activation='relu', normalization=None, dropout_rate=0,
                         dropout_variant=None):
    """RecurrentConvBlock.call -- blocks.py:380-398; dropout layers with dim=3 (:371-374)."""
    y = dropout(p, x5, dropout_rate, dropout_variant)
    y = convlstm2d(p, name + '/convlstm1', y, filters, 5)
    if normalization is not None:
        y = normalize_5d(p, name + '/norm1', y, normalization)
    y = act(y, activation)
    y = dropout(p, y, dropout_rate, dropout_variant)
    y = convlstm2d(p, name + '/convlstm2', y, filters, 3)
    if normalization is not None:
        y = normalize_5d(p, name + '/norm2', y, normalization)
    return act(y, activation)


def pad_concat(t1, t2):
    """PadConcat.call -- blocks.py:629-656 (zero-pad bottom/right of the smaller one)."""
    y1, x1, y2, x2 = t1.shape[2], t1.shape[3], t2.shape[2], t2.shape[3]
    if y2 < y1:
        t2 = F.pad(t2, (0, 0, 0, y1 - y2))
    elif y2 > y1:
        t1 = F.pad(t1, (0, 0, 0, y2 - y1))
    if x2 < x1:
        t2 = F.pad(t2, (0, x1 - x2, 0, 0))
    elif x2 > x1:
        t1 = F.pad(t1, (0, x2 - x1, 0, 0))
    return torch.cat([t1, t2], dim=1)


# ----------------------------------------------------------------------------------------------
# model graphs
# ----------------------------------------------------------------------------------------------
def _nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _tail(p, x, s_in, init_n_filters, n_filters_aux, n_channels_out, activation,
          output_activation, localcon_layer, aux_name='ConvBlock_aux', normalization=None, convnext=False,
          dropout_rate=0):
    """Shared output module: LCB, aux branch, TransitionLast, two ConvBlocks
    -- sp_postups.py:184-212, sp_preups.py:155-183,291-309."""
    if localcon_layer:
        lws = localized_conv_block(p, 'LocalizedConvBlock', x, 2)
        x = torch.cat([x, lws], dim=1)
    if s_in is not None:
        if convnext:                            # sp_postups.py:191-195
            s = convnext_block(p, 'ConvNextBlock_aux', s_in, n_filters_aux, activation, normalization,
                               use_1x1conv=True)
        else:
            s = conv_block(p, aux_name, s_in, n_filters_aux, activation=activation, normalization=normalization)
        x = torch.cat([x, s], dim=1)
    ks = 7 if convnext else 3                   # sp_postups.py:121,133: `ks` of the backbone branch
    x = transition_block(p, 'TransitionLast', x, init_n_filters)   # default relu (App. B #9)
    x = conv_block(p, 'ConvBlock_tail', x, init_n_filters, activation=None, attention=True,
                   normalization=normalization, ks1=ks, ks2=ks, dropout_rate=dropout_rate)   # no variant: :206-208
    x = conv_block(p, 'ConvBlock_out', x, n_channels_out, activation=output_activation,
                   normalization=normalization, ks1=ks, ks2=ks)
    return x


def _backbone(p, x_in, backbone_block, n_filters, n_blocks, attention, activation, normalization=None,
              dropout_rate=0, dropout_variant=None):
    """Backbone section shared by net_postupsampling / net_pin -- sp_postups.py:132-168,
    sp_preups.py:116-151."""
    init_n_filters = n_filters
    if backbone_block == 'convnext':        # sp_postups.py:120-131, sp_preups.py:104-115
        x = b = _conv(p, 'stem', x_in, n_filters, k=7)
        for i in range(n_blocks):
            n_filters = init_n_filters * (i + 1)
            b = convnext_block(p, 'ConvNextBlock' + str(i + 1), b, n_filters, activation, normalization,
                               use_1x1conv=(i != 0))
        x = transition_block(p, 'TransitionSkip', x, n_filters, activation)
        return x + b, n_filters
    x = b = _conv(p, 'stem', x_in, n_filters)
    for i in range(n_blocks):
        n_filters = init_n_filters * (i + 1)
        if backbone_block == 'convnet':
            b = conv_block(p, 'ConvBlock' + str(i + 1), b, n_filters, activation, attention,
                           normalization=normalization, dropout_rate=dropout_rate, dropout_variant=dropout_variant)
        elif backbone_block == 'resnet':
            b = residual_block(p, 'ResidualBlock' + str(i + 1), b, n_filters, activation,
                               attention, use_1x1conv=(i != 0), normalization=normalization,
                               dropout_rate=dropout_rate, dropout_variant=dropout_variant)
        elif backbone_block == 'densenet':
            b = dense_block(p, 'DenseBlock' + str(i + 1), b, n_filters, activation, attention,
                            normalization=normalization, dropout_rate=dropout_rate, dropout_variant=dropout_variant)
            b = transition_block(p, 'Transition' + str(i + 1), b, b.shape[1] // 2)
        else:
            raise NotImplementedError(backbone_block)
    b = act(_conv(p, 'backbone_last', b, n_filters), activation)
    b = dropout(p, b, dropout_rate, dropout_variant)                # sp_postups.py:158
    if backbone_block == 'convnet':
        x = b
    elif backbone_block == 'resnet':
        x = transition_block(p, 'TransitionSkip', x, n_filters, activation)
        x = x + b
    elif backbone_block == 'densenet':
        x = torch.cat([x, b], dim=1)
        x = transition_block(p, 'TransitionBackboneLast', x, n_filters, activation)
    return x, n_filters


def net_postupsampling(p, inputs, backbone_block, upsampling, scale, n_channels_out=1,
                       n_filters=8, n_blocks=6, attention=False, activation='relu',
                       output_activation=None, localcon_layer=False, normalization=None, dropout_rate=0,
                       dropout_variant=None, rc_interpolation='bilinear'):
    """net_postupsampling -- sp_postups.py:14-217.  inputs: [x_lr NHWC] or [x_lr, s_hr]."""
    x_in = _nchw(inputs[0])
    s_in = _nchw(inputs[1]) if len(inputs) > 1 else None
    init_n_filters = n_filters
    x, n_filters = _backbone(p, x_in, backbone_block, n_filters, n_blocks, attention, activation, normalization,
                             dropout_rate, dropout_variant)
    if upsampling == 'spc':
        x = subpixel_block(p, 'SubpixelConvolution', x, scale, n_filters)
    elif upsampling == 'rc':
        x = resize_conv_block(p, 'ResizeConvolution', x, scale, n_filters, rc_interpolation)
    elif upsampling == 'dc':
        x = transition_block(p, 'TransitionDC', x, init_n_filters, activation)
        x = deconv_block(p, 'Deconvolution', x, scale, n_filters, activation)
    x = _tail(p, x, s_in, init_n_filters, n_filters, n_channels_out, activation,
              output_activation, localcon_layer, normalization=normalization,
              convnext=(backbone_block == 'convnext'), dropout_rate=dropout_rate)
    return _nhwc(x)


def net_pin(p, inputs, backbone_block, n_channels_out=1, n_filters=8, n_blocks=6,
            attention=False, activation='relu', output_activation=None, localcon_layer=False,
            normalization=None, dropout_rate=0, dropout_variant=None):
    """net_pin -- sp_preups.py:13-189."""
    x_in = _nchw(inputs[0])
    s_in = _nchw(inputs[1]) if len(inputs) > 1 else None
    init_n_filters = n_filters
    x, n_filters = _backbone(p, x_in, backbone_block, n_filters, n_blocks, attention, activation, normalization,
                             dropout_rate, dropout_variant)
    x = _tail(p, x, s_in, init_n_filters, n_filters, n_channels_out, activation,
              output_activation, localcon_layer, normalization=normalization,
              convnext=(backbone_block == 'convnext'), dropout_rate=dropout_rate)
    return _nhwc(x)


def check_nblocks(shape, power):
    """_check_nblocks -- sp_preups.py:318-324."""
    while shape[0] // 2 ** power < 2 or shape[1] // 2 ** power < 2:
        power -= 1
    return power


def unet_pin(p, inputs, n_filters, n_blocks, n_channels_out=1, activation='relu',
             attention=False, decoder_upsampling='rc', output_activation=None, width_cap=256,
             localcon_layer=False, normalization=None, dropout_rate=0, dropout_variant=None,
             rc_interpolation='bilinear'):
    """unet_pin -- sp_preups.py:192-315 (the bottleneck block is never normalised, :266-268)."""
    x = _nchw(inputs[0])
    s_in = _nchw(inputs[1]) if len(inputs) > 1 else None
    n_blocks = check_nblocks((x.shape[2], x.shape[3]), n_blocks)
    init_n_filters = n_filters
    skips, flist = [], []
    for i in range(n_blocks):
        y = conv_block(p, 'EncoderBlock%d' % (i + 1), x, n_filters, activation, attention,
                       normalization=normalization)
        skips.append(y)
        x = maxpool2(y)
        flist.append(n_filters)
        n_filters = min(width_cap, n_filters * 2)
    x = conv_block(p, 'Bottleneck', x, n_filters, activation, dropout_rate=dropout_rate,
                   dropout_variant=dropout_variant)       # encoder blocks get rate 0 (sp_preups.py:255)
    flist = flist[::-1]
    for j, skip in enumerate(reversed(skips)):
        n_filters = flist[j]
        if decoder_upsampling == 'spc':
            x = subpixel_block(p, 'SubpixelConvolution%d' % (j + 1), x, 2, n_filters)
        elif decoder_upsampling == 'rc':
            x = resize_conv_block(p, 'ResizeConvolution%d' % (j + 1), x, 2, n_filters, rc_interpolation)
        elif decoder_upsampling == 'dc':
            x = deconv_block(p, 'Deconvolution%d' % (j + 1), x, 2, n_filters, activation)
        x = pad_concat(x, skip)
        x = conv_block(p, 'DecoderConvBlock%d' % (j + 1), x, n_filters, activation, attention,
                       normalization=normalization)
    x = dropout(p, x, dropout_rate, dropout_variant)          # sp_preups.py:287
    x = _tail(p, x, s_in, init_n_filters, n_filters, n_channels_out, activation,
              output_activation, localcon_layer, normalization=normalization, dropout_rate=dropout_rate)
    return _nhwc(x)


def recnet_postupsampling(p, inputs, backbone_block, upsampling, scale, time_window,
                          n_channels_out=1, n_filters=8, n_blocks=4, attention=False,
                          activation='relu', output_activation=None, localcon_layer=False, normalization=None,
                          dropout_rate=0, dropout_variant=None, rc_interpolation='bilinear'):
    """recnet_postupsampling -- spt_postups.py:12-163.  inputs[0]: (B,T,h,w,C) NTHWC;
    optional inputs[1]: (B,H,W,n_aux).  Output (B,T,H,W,n_channels_out)."""
    x5 = inputs[0].permute(0, 1, 4, 2, 3).contiguous()   # (B,T,C,h,w)
    bsz, t = x5.shape[0], x5.shape[1]
    nz = normalization
    x = b = recurrent_conv_block(p, 'RecurrentConvBlock1', x5, n_filters, activation, nz)
    for i in range(n_blocks):
        b = recurrent_conv_block(p, 'RecurrentConvBlock' + str(i + 2), b, n_filters, activation, nz, dropout_rate,
                                 dropout_variant)
    b = dropout(p, b, dropout_rate, dropout_variant)              # spt_postups.py:113
    if backbone_block == 'convnet':
        x = b
    elif backbone_block == 'resnet':
        x = x + b
    elif backbone_block == 'densenet':
        x = torch.cat([x, b], dim=2)
    n_filters_ups = x.shape[2]
    xf = x.reshape(bsz * t, x.shape[2], x.shape[3], x.shape[4])   # TimeDistributed
    if upsampling == 'spc':
        xf = subpixel_block(p, 'SubpixelConvolution', xf, scale, n_filters_ups)
    elif upsampling == 'rc':
        xf = resize_conv_block(p, 'ResizeConvolution', xf, scale, n_filters_ups, rc_interpolation)
    elif upsampling == 'dc':
        xf = deconv_block(p, 'Deconvolution', xf, scale, n_filters_ups, None)   # App. B #8
    # upsampling == 'pin' (recnet_pin, spt_preups.py:100-118): no upsampler
    if len(inputs) > 1:
        s = conv_block(p, 'ConvBlock_aux', _nchw(inputs[1]), n_filters, activation, attention)
        s = s.unsqueeze(1).expand(bsz, t, *s.shape[1:]).reshape(bsz * t, *s.shape[1:])
        xf = torch.cat([xf, s], dim=1)
    if localcon_layer:
        lws = localized_conv_block(p, 'LocalizedConvBlock', xf, 2)
        xf = torch.cat([xf, lws], dim=1)
    # spt_postups.py:150 halves the channel count; spt_preups.py:133 maps to n_filters
    xf = transition_block(p, 'TransitionLast', xf, n_filters if upsampling == 'pin' else xf.shape[1] // 2)
    # ConvBlock(n_filters, activation=None, attention=True) on a 5-D tensor (App. B #6)
    # (plain Dropout in front of each convolution when dropout_rate > 0 -- no variant passed, :152-153; the masks
    #  are handed over on the 5-D view so that the supplier sees (B,T,C,H,W) everywhere in this network)
    def drop5(v):
        return dropout(p, v.reshape(bsz, t, *v.shape[1:]), dropout_rate).reshape(v.shape)
    y = _conv(p, 'ConvBlock_tail/conv1', drop5(xf), n_filters, bias=nz is None)
    if nz is not None:
        y = normalize(p, 'ConvBlock_tail/norm1', y, nz)
    y = _conv(p, 'ConvBlock_tail/conv2', drop5(y), n_filters, bias=nz is None)
    if nz is not None:
        y = normalize(p, 'ConvBlock_tail/norm2', y, nz)
    y5 = y.reshape(bsz, t, *y.shape[1:])
    y5 = channel_attention_5d(p, 'ConvBlock_tail/att', y5, n_filters)
    y = y5.reshape(bsz * t, *y5.shape[2:])
    y = conv_block(p, 'ConvBlock_out', y, n_channels_out, activation=output_activation, normalization=nz)
    y5 = y.reshape(bsz, t, *y.shape[1:])
    return y5.permute(0, 1, 3, 4, 2).contiguous()


def recnet_pin(p, inputs, backbone_block, time_window, n_channels_out=1, n_filters=8, n_blocks=6, attention=False,
               activation='relu', output_activation=None, localcon_layer=False, normalization=None, dropout_rate=0,
               dropout_variant=None):
    """recnet_pin -- spt_preups.py:12-163: recnet_postupsampling's graph without the upsampler (inputs already on
    the HR grid) and with TransitionLast -> n_filters (:133)."""
    return recnet_postupsampling(p, inputs, backbone_block, 'pin', 1, time_window, n_channels_out=n_channels_out,
                                 n_filters=n_filters, n_blocks=n_blocks, attention=attention, activation=activation,
                                 output_activation=output_activation, localcon_layer=localcon_layer,
                                 normalization=normalization, dropout_rate=dropout_rate,
                                 dropout_variant=dropout_variant)


def residual_discriminator(p, inputs, upsampling, scale, lr_size, n_filters=8, n_res_blocks=4,
                           attention=False, dropout_mask=None, normalization=None, is_spatiotemporal=False,
                           activation='relu'):
    """residual_discriminator -- discriminator.py:11-81.  ResidualBlocks always relu
    (App. B #10).  ``dropout_mask``: (B, 2*n_filters) keep-mask already scaled by 1/(1-0.4)
    (Dropout(0.4) with training=True, cgan.py:599-600); None = inference (identity).
    ``is_spatiotemporal`` (:25-33,42-47,73-74): inputs (B,T,h,w,C) / (B,T,H,W,1) NTHWC; the LR branch opens with
    RecurrentConvBlock(n_filters, activation, normalization='ln'); Keras Conv2D on a 5-D tensor treats the leading
    axes as batch, so every later layer acts per frame on (B*T,C,H,W); GlobalAveragePooling3D pools (T,H,W)."""
    if is_spatiotemporal:
        x5 = inputs[0].permute(0, 1, 4, 2, 3).contiguous()          # (B,T,C,h,w)
        bsz, t = x5.shape[0], x5.shape[1]
        r5 = inputs[1].permute(0, 1, 4, 2, 3).contiguous()
        x_ref = r5.reshape(bsz * t, *r5.shape[2:])
        b5 = recurrent_conv_block(p, 'branch1_recurrent', x5, n_filters, activation, 'ln')
        x1 = b = b5.reshape(bsz * t, *b5.shape[2:])
    else:
        x_in, x_ref = _nchw(inputs[0]), _nchw(inputs[1])
        x1 = b = _conv(p, 'branch1_stem', x_in, n_filters)
    for i in range(n_res_blocks):
        b = residual_block(p, 'ResidualBlock%d_branch1' % (i + 1), b, n_filters, 'relu', attention,
                           normalization=normalization)
    b = _conv(p, 'branch1_last', b, n_filters)
    x1 = x1 + b
    x2 = c = _conv(p, 'branch2_stem', x_ref, n_filters)
    for i in range(n_res_blocks):
        c = residual_block(p, 'ResidualBlock%d_branch2' % (i + 1), c, n_filters, 'relu', attention,
                           normalization=normalization)
    if upsampling in POSTUPSAMPLING_METHODS:
        if scale == 5:
            c = _conv(p, 'branch2_down1', c, n_filters, stride=2, padding='valid')
            x2 = _conv(p, 'branch2_down2', c, n_filters, stride=2, padding='valid')
            x2 = x2[:, :, :-1, :-1]
        elif scale == 4:
            c = _conv(p, 'branch2_down1', c, n_filters, stride=2)
            x2 = _conv(p, 'branch2_down2', c, n_filters, stride=2)
        else:
            x2 = resize_bilinear(c, lr_size[0], lr_size[1])
    else:
        c = _conv(p, 'branch2_last', c, n_filters)
        x2 = x2 + c
    x = torch.cat([x1, x2], dim=1)
    x = residual_block(p, 'ResidualBlock_merged', x, x.shape[1], 'relu', attention, normalization=normalization)
    if is_spatiotemporal:
        x = x.reshape(bsz, t, *x.shape[1:]).mean(dim=(1, 3, 4))     # GlobalAveragePooling3D
    else:
        x = x.mean(dim=(2, 3))
    if dropout_mask is not None:
        x = x * dropout_mask
    w1 = p.get('dense1/kernel', (x.shape[1], 32))
    b1 = p.get('dense1/bias', (32,))
    x = torch.sigmoid(x @ w1 + b1)
    w2 = p.get('dense2/kernel', (32, 1))
    b2 = p.get('dense2/bias', (1,))
    return torch.sigmoid(x @ w2 + b2)


# ----------------------------------------------------------------------------------------------
# losses / optimizer / steps
# ----------------------------------------------------------------------------------------------
def mae(y_true, y_pred):
    """losses.mae -- losses.py:5-11 (Keras MeanAbsoluteError = global mean)."""
    return (y_true - y_pred).abs().mean()


def mse(y_true, y_pred):
    """losses.mse -- losses.py:14-20."""
    return ((y_true - y_pred) ** 2).mean()


# SSIM family -- losses.py:23-147.  The arithmetic is TensorFlow's `tf.image.ssim` / `tf.image.ssim_multiscale`
# (tensorflow/python/ops/image_ops_impl.py, TF 2.6-2.15: `_fspecial_gauss`, `_ssim_helper`, `_ssim_per_channel`,
# `ssim`, `ssim_multiscale`; third-party, not under /root/reference, unpinned -- restated from the published source).
_MSSSIM_POWER_FACTORS = (0.0448, 0.2856, 0.3001, 0.2363)        # losses.py:128 (4 scales, not TF's 5)


def _fspecial_gauss(size=11, sigma=1.5, dtype=torch.float32):
    """softmax over the flattened 2-D grid of -(x^2+y^2)/(2 sigma^2): `_fspecial_gauss`."""
    coords = torch.arange(size, dtype=dtype) - (size - 1) / 2.0
    g = coords ** 2 * (-0.5 / sigma ** 2)
    g = (g[None, :] + g[:, None]).reshape(1, -1)
    return torch.softmax(g, dim=-1).reshape(size, size)


def _ssim_per_channel(img1, img2, max_val, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03):
    """`_ssim_per_channel` + `_ssim_helper` on NHWC tensors: depthwise VALID gaussian filtering of x, y, x*y and
    x^2+y^2; returns (mean over H,W of luminance*cs, mean over H,W of cs), each (B, C)."""
    c = img1.shape[-1]
    kern = _fspecial_gauss(filter_size, filter_sigma, img1.dtype).reshape(1, 1, filter_size, filter_size)
    kern = kern.repeat(c, 1, 1, 1)
    red = lambda t: F.conv2d(_nchw(t), kern, groups=c)
    c1 = (k1 * max_val) ** 2
    c2 = (k2 * max_val) ** 2
    mean0, mean1 = red(img1), red(img2)
    num0 = mean0 * mean1 * 2.0
    den0 = mean0 ** 2 + mean1 ** 2
    luminance = (num0 + c1) / (den0 + c1)
    num1 = red(img1 * img2) * 2.0
    den1 = red(img1 ** 2 + img2 ** 2)
    cs = (num1 - num0 + c2) / (den1 - den0 + c2)
    return (luminance * cs).mean(dim=(2, 3)), cs.mean(dim=(2, 3))


def tf_image_ssim(img1, img2, max_val, **kw):
    """tf.image.ssim: mean over channels of the per-channel SSIM -> (B,)."""
    return _ssim_per_channel(img1, img2, max_val, **kw)[0].mean(dim=-1)


def tf_image_ssim_multiscale(img1, img2, max_val, power_factors=_MSSSIM_POWER_FACTORS, **kw):
    """tf.image.ssim_multiscale: per scale relu(cs) (relu(ssim) at the last), 2x2 average pooling between scales
    (SYMMETRIC padding of odd sizes), weighted geometric mean over scales, mean over channels -> (B,)."""
    imgs = [img1, img2]
    mcs = []
    for k in range(len(power_factors)):
        if k > 0:
            nxt = []
            for t in imgs:
                t = _nchw(t)
                ph, pw = t.shape[2] % 2, t.shape[3] % 2
                if ph or pw:
                    t = F.pad(t, (0, pw, 0, ph), mode='replicate')   # SYMMETRIC pad by one == edge replicate
                nxt.append(_nhwc(F.avg_pool2d(t, 2, 2)))
            imgs = nxt
        ssim_pc, cs = _ssim_per_channel(imgs[0], imgs[1], max_val, **kw)
        mcs.append(torch.relu(cs))
    mcs.pop()
    stack = torch.stack(mcs + [torch.relu(ssim_pc)], dim=-1)
    pf = torch.tensor(power_factors, dtype=img1.dtype)
    return torch.prod(stack ** pf, dim=-1).mean(dim=-1)


def _positive_pair(y_true, y_pred):
    """losses.py:44-54 / 118-128: dynamic range over both tensors, each shifted by its own minimum if negative."""
    maxv = torch.maximum(y_true.max(), y_pred.max())
    minv = torch.minimum(y_true.min(), y_pred.min())
    drange = maxv - minv
    yt = y_true - y_true.min() if y_true.min() < 0 else y_true
    yp = y_pred - y_pred.min() if y_pred.min() < 0 else y_pred
    return yt, yp, drange


def dssim(y_true, y_pred):
    """losses.dssim -- losses.py:27-59."""
    yt, yp, drange = _positive_pair(y_true, y_pred)
    return ((1.0 - tf_image_ssim(yt, yp, drange)) / 2.0).mean()


def msdssim(y_true, y_pred):
    """losses.msdssim -- losses.py:96-131."""
    yt, yp, drange = _positive_pair(y_true, y_pred)
    return ((1.0 - tf_image_ssim_multiscale(yt, yp, drange)) / 2.0).mean()


def dssim_mae(y_true, y_pred):
    """losses.py:62-68."""
    return 0.8 * dssim(y_true, y_pred) + 0.2 * mae(y_true, y_pred)


def dssim_mae_mse(y_true, y_pred):
    """losses.py:71-84."""
    return 0.6 * dssim(y_true, y_pred) + 0.2 * mae(y_true, y_pred) + 0.2 * mse(y_true, y_pred)


def dssim_mse(y_true, y_pred):
    """losses.py:87-93."""
    return 0.8 * dssim(y_true, y_pred) + 0.2 * mse(y_true, y_pred)


def msdssim_mae(y_true, y_pred):
    """losses.py:134-140."""
    return 0.8 * msdssim(y_true, y_pred) + 0.2 * mae(y_true, y_pred)


def msdssim_mae_mse(y_true, y_pred):
    """losses.py:143-151."""
    return 0.6 * msdssim(y_true, y_pred) + 0.2 * mae(y_true, y_pred) + 0.2 * mse(y_true, y_pred)


LOSSES = {'mae': mae, 'mse': mse, 'dssim': dssim, 'dssim_mae': dssim_mae, 'dssim_mse': dssim_mse,
          'dssim_mae_mse': dssim_mae_mse, 'msdssim': msdssim, 'msdssim_mae': msdssim_mae,
          'msdssim_mae_mse': msdssim_mae_mse}


def bce(y_true, p):
    """tf.keras.losses.BinaryCrossentropy(from_logits=False) -- cgan.py:546,567.
    Keras backend: p = clip(p, eps, 1-eps); -mean(y log(p+eps) + (1-y) log(1-p+eps)), eps=1e-7."""
    eps = 1e-7
    p = torch.clamp(p, eps, 1.0 - eps)
    return -(y_true * torch.log(p + eps) + (1.0 - y_true) * torch.log(1.0 - p + eps)).mean()


class TFAdam:
    """tf.keras.optimizers.Adam (beta2 0.999, eps 1e-7 OUTSIDE the bias correction)
    -- supervised.py:353, cgan.py:277-278.
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t * m / (sqrt(v) + eps)."""

    def __init__(self, names, lr=1e-3, beta_1=0.9, beta_2=0.999, eps=1e-7):
        self.lr, self.b1, self.b2, self.eps = lr, beta_1, beta_2, eps
        self.t = 0
        self.m = {n: None for n in names}
        self.v = {n: None for n in names}

    def lr_at(self, step):
        lr = self.lr
        if callable(lr):
            return lr(step)
        return lr

    def apply(self, weights, grads):
        self.t += 1
        lr = self.lr_at(self.t - 1)   # Keras evaluates the schedule at `iterations` before increment
        lr_t = lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        with torch.no_grad():
            for n, w in weights.items():
                g = grads[n]
                if self.m[n] is None:
                    self.m[n] = torch.zeros_like(w)
                    self.v[n] = torch.zeros_like(w)
                self.m[n].mul_(self.b1).add_(g, alpha=1.0 - self.b1)
                self.v[n].mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
                w.sub_(lr_t * self.m[n] / (self.v[n].sqrt() + self.eps))


def piecewise_constant(boundary, v0, v1):
    """PiecewiseConstantDecay([b],[v0,v1]): v0 while step <= b else v1 -- supervised.py:340-346."""
    return lambda step: v0 if step <= boundary else v1


def supervised_step(forward_fn, weights, opt, inputs, target, loss='mae'):
    """One Keras ``fit`` train step: fwd, loss, bwd, Adam -- supervised.py:396-406.
    ``forward_fn(Params, inputs) -> y``.  Returns (loss value, grads dict)."""
    for w in weights.values():
        w.requires_grad_(True)
        w.grad = None
    y = forward_fn(Params(weights), inputs)
    lossv = LOSSES[loss](target, y)
    lossv.backward()
    # (variables without a gradient -- BN moving statistics, DenseBlock's unused norm1 -- are skipped by Keras)
    grads = {n: (w.grad.detach().clone() if w.grad is not None else torch.zeros_like(w)) for n, w in weights.items()}
    for w in weights.values():
        w.requires_grad_(False)
    if opt is not None:
        opt.apply(weights, grads)
    return float(lossv.detach()), grads


def cgan_step(gen_fn, disc_fn, gw, dw, gopt, dopt, lr_array, hr_array, static_array=None,
              loss='mae', mask_real=None, mask_fake=None, lam=100.0):
    """train_step -- cgan.py:575-617.  ``gen_fn(Params, inputs)``, ``disc_fn(Params, inputs,
    dropout_mask)``.  Returns ((gen_total, gen_gan, gen_px, disc), g_grads, d_grads)."""
    for w in list(gw.values()) + list(dw.values()):
        w.requires_grad_(True)
        w.grad = None
    gin = [lr_array, static_array] if static_array is not None else [lr_array]
    gen = gen_fn(Params(gw), gin)
    d_real = disc_fn(Params(dw), [lr_array, hr_array], mask_real)
    d_fake = disc_fn(Params(dw), [lr_array, gen], mask_fake)
    gan = bce(torch.ones_like(d_fake), d_fake)
    px = LOSSES[loss](hr_array, gen)
    g_total = gan + lam * px
    d_loss = bce(torch.ones_like(d_real), d_real) + bce(torch.zeros_like(d_fake), d_fake)
    g_grads = torch.autograd.grad(g_total, list(gw.values()), retain_graph=True)
    d_grads = torch.autograd.grad(d_loss, list(dw.values()))
    g_grads = {n: g.detach().clone() for n, g in zip(gw.keys(), g_grads)}
    d_grads = {n: g.detach().clone() for n, g in zip(dw.keys(), d_grads)}
    for w in list(gw.values()) + list(dw.values()):
        w.requires_grad_(False)
    if gopt is not None:
        gopt.apply(gw, g_grads)
        dopt.apply(dw, d_grads)
    return (float(g_total.detach()), float(gan.detach()), float(px.detach()),
            float(d_loss.detach())), g_grads, d_grads


# ----------------------------------------------------------------------------------------------
# weight init (Keras defaults), seeded numpy -> shared by oracle and CUDA backends in tests
# ----------------------------------------------------------------------------------------------
def glorot_uniform(rng, shape):
    """Keras glorot_uniform: U(+-sqrt(6/(fan_in+fan_out))); conv fans use the receptive field."""
    if len(shape) == 1:
        fan_in = fan_out = shape[0]
    elif len(shape) == 2:
        fan_in, fan_out = shape
    else:
        rf = int(np.prod(shape[:-2]))
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def init_weights(spec, seed=0, bias_scale=0.0):
    """Seeded weights for a spec {name: shape}.  Kernels glorot-uniform; biases zero (Keras) or,
    with ``bias_scale``>0, small random values so bias paths are exercised by parity tests."""
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for name, shape in spec.items():
        if name.endswith(('/gamma', '/moving_variance')):
            # Keras: ones; with bias_scale > 0 perturbed so that the scale paths are exercised
            a = 1.0 + (bias_scale * rng.uniform(-1, 1, size=shape) if bias_scale > 0 else np.zeros(shape))
            out[name] = torch.from_numpy(a.astype(np.float32))
        elif name.endswith(('/beta', '/moving_mean')):
            a = bias_scale * rng.standard_normal(shape) if bias_scale > 0 else np.zeros(shape)
            out[name] = torch.from_numpy(a.astype(np.float32))
        elif name.endswith('/bias'):
            if bias_scale > 0:
                out[name] = torch.from_numpy((bias_scale * rng.standard_normal(shape)).astype(np.float32))
            else:
                out[name] = torch.zeros(shape, dtype=torch.float32)
        elif name.endswith('localconv/kernel'):
            n = int(np.prod(shape))
            lim = math.sqrt(3.0 / n) if bias_scale == 0 else 0.5
            out[name] = torch.from_numpy(rng.uniform(-lim, lim, size=shape).astype(np.float32))
        else:
            out[name] = torch.from_numpy(glorot_uniform(rng, shape))
    return out
