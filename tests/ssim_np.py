"""numpy fp64 statements used by the SSIM-loss tests (test infrastructure):

* ``ssim_direct``: SSIM of one image pair straight from the definition (explicit 11x11 windows, Wang et al. 2004
  with tf.image.ssim's constants) -- an independent check of oracle/torch_ref.py's convolution-based restatement;
* ``loss_and_grad``: the closed-form backward the CUDA kernels implement (dl4ds_b200/csrc/ssim.cu), checked on
  the CPU against torch autograd of the oracle so that the derivation itself is pinned without a GPU.
"""
import numpy as np
from scipy.signal import convolve2d, correlate2d

K1, K2 = 0.01, 0.03
POWER_FACTORS = (0.0448, 0.2856, 0.3001, 0.2363)     # losses.py:128


def gauss2d(size=11, sigma=1.5):
    c = np.arange(size) - (size - 1) / 2.0
    g = np.exp(-(c[:, None] ** 2 + c[None, :] ** 2) / (2.0 * sigma ** 2))
    return g / g.sum()


def ssim_direct(x, y, L):
    """mean SSIM of two 2-D arrays, window by window."""
    w = gauss2d()
    c1, c2 = (K1 * L) ** 2, (K2 * L) ** 2
    H, W = x.shape
    vals = []
    for i in range(H - 10):
        for j in range(W - 10):
            a, b = x[i:i + 11, j:j + 11], y[i:i + 11, j:j + 11]
            mx, my = (w * a).sum(), (w * b).sum()
            vx, vy = (w * a * a).sum() - mx * mx, (w * b * b).sum() - my * my
            cxy = (w * a * b).sum() - mx * my
            vals.append((2 * mx * my + c1) * (2 * cxy + c2) / ((mx * mx + my * my + c1) * (vx + vy + c2)))
    return float(np.mean(vals))


def _filt(a):
    return correlate2d(a, gauss2d(), mode='valid')


def _filt_t(g):
    return convolve2d(g, gauss2d(), mode='full')


def _maps(x, y, C1, C2, cs_only):
    mx, my, exy, e2 = _filt(x), _filt(y), _filt(x * y), _filt(x * x + y * y)
    B1 = mx * mx + my * my + C1
    B2 = e2 - mx * mx - my * my + C2
    lum = (2 * mx * my + C1) / B1
    cs = (2 * exy - 2 * mx * my + C2) / B2
    dlum_dmx = (2 * my - lum * 2 * mx) / B1
    dlum_dC1 = (1 - lum) / B1
    if cs_only:
        lum, dlum_dmx, dlum_dC1 = np.ones_like(cs), np.zeros_like(cs), np.zeros_like(cs)
    dcs_dmx = (2 * mx * cs - 2 * my) / B2
    return (lum * cs, cs * dlum_dmx + lum * dcs_dmx, -lum * cs / B2, 2 * lum / B2, cs * dlum_dC1,
            lum * (1 - cs) / B2)


def _pool(a):
    """ssim_multiscale's downsampling: SYMMETRIC pad of odd sizes by one (= repeat the edge), 2x2 mean."""
    a = np.pad(a, ((0, a.shape[0] % 2), (0, a.shape[1] % 2)), mode='edge')
    return a.reshape(a.shape[0] // 2, 2, a.shape[1] // 2, 2).mean(axis=(1, 3))


def _unpool(g, shape):
    """adjoint of _pool: a quarter of the coarse gradient to each of the four sources; a mirrored edge row / column
    is its own neighbour and collects twice."""
    up = np.repeat(np.repeat(g, 2, 0), 2, 1) / 4.0
    out = up[:shape[0], :shape[1]].copy()
    if shape[0] % 2:
        out[-1, :] += up[shape[0], :shape[1]]
    if shape[1] % 2:
        out[:, -1] += up[:shape[0], shape[1]]
    if shape[0] % 2 and shape[1] % 2:
        out[-1, -1] += up[shape[0], shape[1]]
    return out


def loss_and_grad(y_true, y_pred, multiscale):
    """y_true, y_pred (B,H,W) fp64 -> (dssim or msdssim, d loss / d y_pred), kernel algorithm of ssim.cu."""
    B = y_true.shape[0]
    maxt, mint, maxp, minp = y_true.max(), y_true.min(), y_pred.max(), y_pred.min()
    L = max(maxt, maxp) - min(mint, minp)
    st = mint if mint < 0 else 0.0
    sp = minp if minp < 0 else 0.0
    C1, C2 = (K1 * L) ** 2, (K2 * L) ** 2
    pf = POWER_FACTORS if multiscale else (1.0,)
    nS = len(pf)
    dy = np.zeros_like(y_pred)
    loss, dL = 0.0, 0.0
    for b in range(B):
        xs, ys = [y_pred[b] - sp], [y_true[b] - st]
        for _ in range(1, nS):
            xs.append(_pool(xs[-1]))
            ys.append(_pool(ys[-1]))
        maps = [_maps(xs[j], ys[j], C1, C2, j < nS - 1) for j in range(nS)]
        raw = [m[0].mean() for m in maps]
        v = raw if nS == 1 else [max(r, 0.0) for r in raw]
        msv = v[0] if nS == 1 else float(np.prod([v[j] ** pf[j] for j in range(nS)]))
        loss += (1 - msv) / 2 / B
        g = None
        for j in reversed(range(nS)):
            dms = 1.0 if nS == 1 else (pf[j] * msv / v[j] if v[j] > 0 else 0.0)
            coef = (-0.5 / B) * dms / maps[j][0].size
            gj = coef * (_filt_t(maps[j][1]) + 2 * xs[j] * _filt_t(maps[j][2]) + ys[j] * _filt_t(maps[j][3]))
            if g is not None:
                gj = gj + _unpool(g, gj.shape)
            g = gj
            dL += coef * (maps[j][4].sum() * 2 * K1 * K1 * L + maps[j][5].sum() * 2 * K2 * K2 * L)
        dy[b] = g
    sum_dx = dy.sum()
    ip = np.unravel_index(np.argmax(y_pred), y_pred.shape)
    iq = np.unravel_index(np.argmin(y_pred), y_pred.shape)
    if maxp > maxt:
        dy[ip] += dL
    if minp < mint:
        dy[iq] -= dL
    if minp < 0:
        dy[iq] -= sum_dx
    return loss, dy
