"""Numerical feasibility of the round-2 plan: a 3-term fp16 operand split with
per-tensor power-of-two scales against today's 3-term TF32 split, on GEMMs shaped like the headline layers
(K = 3*3*48 = 432) with forward-like operands (post-ReLU activations x glorot weights) and backward-like ones
(gradients of size ~1e-6).  Products are exact in fp32 for both formats; accumulation is emulated in fp32.
Pure numpy; prints max |err| / max |y| against fp64.   usage: python scratch/f16x3_study.py"""
import numpy as np

rng = np.random.default_rng(0)


def tf32_trunc(x):
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32_rna(x):
    u = x.astype(np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def split_tf32(x):
    hi = tf32_trunc(x)                      # the hardware truncates the raw fp32 operand
    lo = tf32_rna((x.astype(np.float32) - hi))
    return hi, lo


def pow2_scale(x, top=2.0 ** 14):
    m = float(np.abs(x).max())
    return 2.0 ** np.floor(np.log2(top / m)) if m > 0 else 1.0


def split_f16(x, s):
    xs = x.astype(np.float32) * np.float32(s)
    hi = xs.astype(np.float16)
    lo = (xs - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def gemm32(a, b):
    """fp32 accumulation of exact products, in K chunks of 16 (the order inside a tensor-core MMA is not specified;
    the chunked sum is representative)."""
    acc = np.zeros((a.shape[0], b.shape[1]), np.float32)
    for k in range(0, a.shape[1], 16):
        acc += (a[:, k:k + 16].astype(np.float64) @ b[k:k + 16].astype(np.float64)).astype(np.float32)
    return acc


def study(label, a, w):
    ref = a.astype(np.float64) @ w.astype(np.float64)
    scale = np.abs(ref).max()
    ah, al = split_tf32(a)
    wh, wl = split_tf32(w)
    y_tf32 = gemm32(tf32_rna(a), tf32_rna(w))
    y_tf32x3 = gemm32(ah, wh) + gemm32(ah, wl) + gemm32(al, wh)
    sa, sw = pow2_scale(a), pow2_scale(w)
    fh, fl = split_f16(a, sa)
    gh, gl = split_f16(w, sw)
    y_f16x3 = (gemm32(fh, gh) + gemm32(fh, gl) + gemm32(fl, gh)) / np.float32(sa * sw)
    y_f16 = gemm32(fh, gh) / np.float32(sa * sw)
    y_fp32 = gemm32(a.astype(np.float32), w.astype(np.float32))
    e = lambda y: np.abs(y.astype(np.float64) - ref).max() / scale
    print('%-34s fp32 %.1e | tf32 %.1e | tf32x3 %.1e | f16 %.1e | f16x3 %.1e   (scales 2^%d, 2^%d)' % (
        label, e(y_fp32), e(y_tf32), e(y_tf32x3), e(y_f16), e(y_f16x3), int(np.log2(sa)), int(np.log2(sw))))


M, K, N = 2048, 432, 48
w = rng.uniform(-0.1, 0.1, (K, N)).astype(np.float32)
study('forward: relu(N(0,1)) x glorot', np.maximum(rng.standard_normal((M, K)), 0).astype(np.float32), w)
study('forward: N(0,1) x glorot', rng.standard_normal((M, K)).astype(np.float32), w)
study('forward: heavy tail (x^3) x glorot', (rng.standard_normal((M, K)) ** 3).astype(np.float32), w)
study('dgrad: 1e-6 N(0,1) x glorot', (1e-6 * rng.standard_normal((M, K))).astype(np.float32), w)
study('dgrad: sparse 1e-6 (90% zeros)', (1e-6 * rng.standard_normal((M, K)) * (rng.random((M, K)) > 0.9)).astype(np.float32), w)
# wgrad: pixels are the K dimension (long sums)
a = np.maximum(rng.standard_normal((432, 16384)), 0).astype(np.float32)
g = (1e-6 * rng.standard_normal((16384, 48))).astype(np.float32)
study('wgrad: K = 16384 pixels', a, g)
