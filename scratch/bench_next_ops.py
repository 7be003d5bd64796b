"""Device timings of the SURVEY 8f row-3 kernels (SSIM losses, batch / layer norm, depthwise 7x7, GELU) at the
headline tensor sizes, with the algorithmic HBM bytes of each call -> achieved GB/s against the measured HBM peak.
CUDA events on the launching stream, warm-up, 20 repeats, a 256 MB L2 flush between repeats.
Usage: python scratch/bench_next_ops.py > profiles/<round>_next_ops.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dl4ds_b200 import _lib, losses          # noqa: E402
from dl4ds_b200.engine import _PF_HOST, _stream   # noqa: E402

dev = torch.device('cuda:0')
peaks = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))
HBM = float(peaks.get('hbm_gbs_sustained', peaks.get('hbm_gbs', 6550.0))) if isinstance(peaks, dict) else 6550.0
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e3     # us


rows = []


def row(name, us, nbytes, note):
    rows.append({'op': name, 'us': round(us, 2), 'algorithmic_MB': round(nbytes / 1e6, 2),
                 'GBps': round(nbytes / us / 1e3, 1), 'frac_hbm_peak': round(nbytes / us / 1e3 / HBM, 3), 'note': note})


# ---- SSIM losses on the headline batch: (64,128,128,1)
B, H, W, C = 64, 128, 128, 1
yp = torch.randn(B, H, W, C, device=dev)
yt = yp + 0.3 * torch.randn_like(yp)
n = yp.numel()
for name, ns in (('dssim', 1), ('msdssim', 4)):
    nws = _lib.load().dl4ds_ssim_loss_workspace_floats(B, H, W, C, ns)
    ws = torch.empty(nws, dtype=torch.float32, device=dev)
    out = torch.zeros(1, device=dev)
    dy = torch.empty_like(yp)
    f = lambda: _lib.call('dl4ds_ssim_loss', yp.data_ptr(), yt.data_ptr(), B, H, W, C, ns, _PF_HOST, 1.0,
                          out.data_ptr(), dy.data_ptr(), 0, ws.data_ptr(), _stream())
    # algorithmic: range pass reads 2 tensors; maps pass reads 2, writes 3 maps; backward reads 2 + 3 maps, writes 1
    scale_sum = sum(0.25 ** j for j in range(ns))
    nbytes = 4 * n * (2 + scale_sum * (2 + 3 + 2 + 3 + 1))
    row('ssim_loss[%s] fwd+bwd (64,128,128,1)' % name, timeit(f), nbytes,
        '%d kernels behind one entry point' % (2 + ns + 2 * (ns - 1) + 1 + ns + 1))

# ---- normalisation
for shape in ((64, 32, 32, 48), (64, 128, 128, 8)):
    x = torch.randn(shape, device=dev)
    y = torch.empty_like(x)
    dyt = torch.randn_like(x)
    dx = torch.empty_like(x)
    Cc = shape[3]
    npix = x.numel() // Cc
    g, b = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
    dg, db = torch.zeros(Cc, device=dev), torch.zeros(Cc, device=dev)
    st = torch.zeros(4 * Cc, device=dev)
    mm, mv = torch.zeros(Cc, device=dev), torch.ones(Cc, device=dev)
    nb = 4 * x.numel()

    def bn_f():
        _lib.call('dl4ds_batchnorm_stats', x.data_ptr(), Cc, npix, Cc, st.data_ptr(), st.data_ptr() + 4 * Cc,
                  mm.data_ptr(), mv.data_ptr(), 0.99, st.data_ptr() + 8 * Cc, _stream())
        _lib.call('dl4ds_norm_apply', x.data_ptr(), Cc, st.data_ptr(), st.data_ptr() + 4 * Cc, g.data_ptr(),
                  b.data_ptr(), 1e-3, y.data_ptr(), Cc, npix, Cc, 1, _stream())
    row('batchnorm+relu fwd %s' % (shape,), timeit(bn_f), nb * 4, 'reads x three times (mean, variance, apply), writes y')

    def bn_b():
        _lib.call('dl4ds_batchnorm_bwd', dyt.data_ptr(), Cc, x.data_ptr(), Cc, y.data_ptr(), Cc, st.data_ptr(),
                  st.data_ptr() + 4 * Cc, g.data_ptr(), 1e-3, dx.data_ptr(), Cc, dg.data_ptr(), db.data_ptr(),
                  st.data_ptr() + 8 * Cc, npix, Cc, 1, _stream())
    row('batchnorm+relu bwd %s' % (shape,), timeit(bn_b), nb * 7, 'two passes over (dy, x, y), writes dx')
    ln_f = lambda: _lib.call('dl4ds_layernorm_fwd', x.data_ptr(), Cc, g.data_ptr(), b.data_ptr(), 1e-3, y.data_ptr(),
                             Cc, npix, Cc, 1, _stream())
    row('layernorm+relu fwd %s' % (shape,), timeit(ln_f), nb * 2, 'reads x, writes y')
    ln_b = lambda: _lib.call('dl4ds_layernorm_bwd', dyt.data_ptr(), Cc, x.data_ptr(), Cc, y.data_ptr(), Cc,
                             g.data_ptr(), 1e-3, dx.data_ptr(), Cc, dg.data_ptr(), db.data_ptr(), npix, Cc, 1,
                             _stream())
    row('layernorm+relu bwd %s' % (shape,), timeit(ln_b), nb * 4, 'reads dy, x, y, writes dx')

# ---- depthwise 7x7 + GELU
x = torch.randn(64, 32, 32, 48, device=dev)
y = torch.empty_like(x)
w = torch.randn(7, 7, 48, 1, device=dev)
bias = torch.zeros(48, device=dev)
dw = torch.zeros_like(w)
nb = 4 * x.numel()
row('depthwise7x7 fwd (64,32,32,48)', timeit(lambda: _lib.call(
    'dl4ds_depthwise_conv_fwd', x.data_ptr(), 48, w.data_ptr(), bias.data_ptr(), y.data_ptr(), 48, 64, 32, 32, 48, 7,
    0, 0, _stream())), nb * 2, '49 MACs per element')
row('depthwise7x7 wgrad (64,32,32,48)', timeit(lambda: _lib.call(
    'dl4ds_depthwise_conv_wgrad', x.data_ptr(), 48, y.data_ptr(), 48, dw.data_ptr(), 64, 32, 32, 48, 7, _stream())),
    nb * 2, '49 MACs per element, register accumulators')
x4 = torch.randn(64, 32, 32, 192, device=dev)
y4 = torch.empty_like(x4)
row('gelu fwd (64,32,32,192)', timeit(lambda: _lib.call('dl4ds_gelu_fwd', x4.data_ptr(), y4.data_ptr(), x4.numel(),
                                                        _stream())), 8 * x4.numel(), '')
print(json.dumps({'hbm_peak_gbs': HBM, 'l2_policy': '256 MB flush buffer written between repeats', 'rows': rows},
                 indent=1))
