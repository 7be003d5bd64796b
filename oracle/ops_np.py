"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- numpy-fp64 direct-definition ops.

An independent statement (explicit index arithmetic, NHWC, no library conv) of the TF/Keras 2.x
op semantics listed in SURVEY.md App. A.  Used on tiny shapes to cross-check
``oracle/torch_ref.py`` so a shared misreading of a definition cannot hide in both the torch
restatement and the CUDA kernels.  PARITY UNPINNED (the reference has no op-level tests).
"""
import numpy as np


def same_pads(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2


def conv2d(x, w, b=None, stride=1, padding='same'):
    """Keras Conv2D, NHWC x HWIO (reference call sites: blocks.py:49-61)."""
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    n, h, wd, cin = x.shape
    kh, kw, _, cout = w.shape
    if padding == 'same':
        ho, pt = same_pads(h, kh, stride)
        wo, pl = same_pads(wd, kw, stride)
    else:
        ho, pt = (h - kh) // stride + 1, 0
        wo, pl = (wd - kw) // stride + 1, 0
    y = np.zeros((n, ho, wo, cout))
    for oy in range(ho):
        for ox in range(wo):
            for ky in range(kh):
                iy = oy * stride + ky - pt
                if iy < 0 or iy >= h:
                    continue
                for kx in range(kw):
                    ix = ox * stride + kx - pl
                    if ix < 0 or ix >= wd:
                        continue
                    y[:, oy, ox, :] += x[:, iy, ix, :] @ w[ky, kx]
    if b is not None:
        y += np.asarray(b, np.float64)
    return y


def conv2d_transpose_same(x, w, stride):
    """Keras Conv2DTranspose(padding='same', use_bias=False), kernel (kh,kw,Cout,Cin)
    (blocks.py:508-516): scatter form of the SAME-conv input gradient."""
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    n, h, wd, cin = x.shape
    kh, kw, cout, _ = w.shape
    ho, wo = h * stride, wd * stride
    _, pt = same_pads(ho, kh, stride)
    _, pl = same_pads(wo, kw, stride)
    y = np.zeros((n, ho, wo, cout))
    for iy in range(h):
        for ix in range(wd):
            for ky in range(kh):
                oy = iy * stride + ky - pt
                if oy < 0 or oy >= ho:
                    continue
                for kx in range(kw):
                    ox = ix * stride + kx - pl
                    if ox < 0 or ox >= wo:
                        continue
                    y[:, oy, ox, :] += x[:, iy, ix, :] @ w[ky, kx].T
    return y


def depth_to_space(x, r):
    """tf.nn.depth_to_space, NHWC DCR (blocks.py:427)."""
    n, h, w, ch = x.shape
    c = ch // (r * r)
    y = np.zeros((n, h * r, w * r, c), x.dtype)
    for i in range(r):
        for j in range(r):
            y[:, i::r, j::r, :] = x[:, :, :, (i * r + j) * c:(i * r + j + 1) * c]
    return y


def resize_bilinear(x, ho, wo):
    """tf.image.resize bilinear, half-pixel centers, no antialias (blocks.py:489)."""
    x = np.asarray(x, np.float64)
    n, h, w, c = x.shape
    y = np.zeros((n, ho, wo, c))
    for oy in range(ho):
        sy = (oy + 0.5) * h / ho - 0.5
        y0 = int(np.floor(sy))
        fy = sy - y0
        ya, yb = min(max(y0, 0), h - 1), min(max(y0 + 1, 0), h - 1)
        for ox in range(wo):
            sx = (ox + 0.5) * w / wo - 0.5
            x0 = int(np.floor(sx))
            fx = sx - x0
            xa, xb = min(max(x0, 0), w - 1), min(max(x0 + 1, 0), w - 1)
            top = x[:, ya, xa] * (1 - fx) + x[:, ya, xb] * fx
            bot = x[:, yb, xa] * (1 - fx) + x[:, yb, xb] * fx
            y[:, oy, ox] = top * (1 - fy) + bot * fy
    return y


def maxpool2(x):
    n, h, w, c = x.shape
    x = x[:, :h // 2 * 2, :w // 2 * 2]
    return x.reshape(n, h // 2, 2, w // 2, 2, c).max(axis=(2, 4))


def local_conv1x1(x, w, b):
    """LocallyConnected2D 1x1 (blocks.py:322-328): W[H,W,Cin,F], b[H,W,F]."""
    return np.einsum('nhwc,hwco->nhwo', np.asarray(x, np.float64), np.asarray(w, np.float64)) + b


def channel_attention(x, w1, b1, w2, b2):
    """ChannelAttention2D (blocks.py:585-593); w1 (1,1,C,C/r), w2 (1,1,C/r,C)."""
    x = np.asarray(x, np.float64)
    y = x.mean(axis=(1, 2))
    y = np.maximum(y @ w1[0, 0] + b1, 0.0)
    y = 1.0 / (1.0 + np.exp(-(y @ w2[0, 0] + b2)))
    return x * y[:, None, None, :]


def hard_sigmoid(x):
    return np.clip(0.2 * x + 0.5, 0.0, 1.0)


def convlstm2d(x, wx, wh, b):
    """ConvLSTM2D return_sequences (blocks.py:350-355); x (B,T,H,W,C)."""
    bsz, t, h, w, _ = x.shape
    f = wh.shape[2]
    hs = np.zeros((bsz, h, w, f))
    cs = np.zeros((bsz, h, w, f))
    out = []
    for ti in range(t):
        z = conv2d(x[:, ti], wx, b) + conv2d(hs, wh)
        i, fg, g, o = z[..., :f], z[..., f:2 * f], z[..., 2 * f:3 * f], z[..., 3 * f:]
        cs = hard_sigmoid(fg) * cs + hard_sigmoid(i) * np.tanh(g)
        hs = hard_sigmoid(o) * np.tanh(cs)
        out.append(hs)
    return np.stack(out, axis=1)


def block_mean(x, s):
    """cv2.resize(INTER_AREA) at integer down-factor s == s x s block mean (utils.py:376-384)."""
    n, h, w, c = x.shape
    return x.reshape(n, h // s, s, w // s, s, c).mean(axis=(2, 4))


def bce(y, p, eps=1e-7):
    p = np.clip(p, eps, 1 - eps)
    return -np.mean(y * np.log(p + eps) + (1 - y) * np.log(1 - p + eps))


def adam_step(theta, g, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-7):
    """tf.keras Adam (supervised.py:353): returns (theta, m, v) after step t (1-based)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    return theta - lr_t * m / (np.sqrt(v) + eps), m, v
