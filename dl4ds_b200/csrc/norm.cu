// BatchNormalization / LayerNormalization (+ fused activation) of the conv blocks -- blocks.py:63-71,94-101,
// 165-176,216-224,263-272,298-305 (Keras defaults: axis -1, epsilon 1e-3, BN momentum 0.99, gamma 1 / beta 0).
//
// Thread mapping shared by every kernel here: a group of G lanes (G = power of two <= 32) owns one pixel, lane l
// holds channels l, l+G, ... (at most kMaxChunk of them: C <= 8 G), so a warp reads 32 consecutive floats of the
// NHWC tensor and the per-pixel (layer norm) or per-channel (batch norm) sums stay in registers.  All HBM-bound:
// layer norm reads x once per pass; batch norm needs the batch statistics first (two passes: mean, then centred
// second moment) and, backward, two per-channel sums before the element-wise pass.
#include <algorithm>
#include <initializer_list>

#include "common.cuh"

namespace dl4ds {
namespace {

constexpr int kMaxChunk = 8;
constexpr int kThreads = 256;

__device__ __forceinline__ float group_sum(float v, int G) {
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

int group_width(int C) {
    int g = 1;
    while (g < C && g < 32) g <<= 1;
    return g;
}

// per-channel partial sums held in registers -> shared -> global (atomicAdd)
template <int NACC>
__device__ __forceinline__ void flush_channel_sums(const float (&acc)[NACC][kMaxChunk], int lane, int G, int C,
                                                   float* sh /* [NACC][256] */, float* const (&dst)[NACC]) {
    for (int i = threadIdx.x; i < NACC * 256; i += kThreads) sh[i] = 0.0f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kMaxChunk; ++k) {
        const int c = lane + k * G;
        if (c < C) {
#pragma unroll
            for (int a = 0; a < NACC; ++a) atomicAdd(sh + a * 256 + c, acc[a][k]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NACC * 256; i += kThreads) {
        const int a = i / 256, c = i % 256;
        if (c < C) atomicAdd(dst[a] + c, sh[i]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// batch norm
// ---------------------------------------------------------------------------------------------------------
// pass 0: sums[c] += sum_p x;  pass 1: sums[C + c] += sum_p (x - sums[c]/M)^2
__global__ void __launch_bounds__(kThreads) bn_stats_kernel(const float* __restrict__ x, int x_ld, int64_t n_pix,
                                                            int C, int G, float* __restrict__ sums, int pass) {
    __shared__ float sh[256];
    const int lane = threadIdx.x % G, row = threadIdx.x / G, rows = kThreads / G;
    float acc[1][kMaxChunk];
    float mu[kMaxChunk];
#pragma unroll
    for (int k = 0; k < kMaxChunk; ++k) {
        acc[0][k] = 0.0f;
        const int c = lane + k * G;
        mu[k] = (pass == 1 && c < C) ? sums[c] / (float)n_pix : 0.0f;
    }
    for (int64_t p = (int64_t)blockIdx.x * rows + row; p < n_pix; p += (int64_t)gridDim.x * rows) {
#pragma unroll
        for (int k = 0; k < kMaxChunk; ++k) {
            const int c = lane + k * G;
            if (c < C) {
                const float d = __ldg(x + p * x_ld + c) - mu[k];
                acc[0][k] += pass == 0 ? d : d * d;
            }
        }
    }
    float* const dst[1] = {sums + (pass == 0 ? 0 : C)};
    flush_channel_sums<1>(acc, lane, G, C, sh, dst);
}

// sums -> (mean, biased variance); moving statistics (Keras fused BN: the moving variance takes the unbiased
// batch variance)
__global__ void bn_finalize_kernel(const float* __restrict__ sums, int C, float M, float* __restrict__ mean,
                                   float* __restrict__ var, float* __restrict__ moving_mean,
                                   float* __restrict__ moving_var, float momentum) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float m = sums[c] / M, v = sums[C + c] / M;
    mean[c] = m;
    var[c] = v;
    if (moving_mean) {
        const float unbiased = M > 1.0f ? v * (M / (M - 1.0f)) : v;
        moving_mean[c] = moving_mean[c] * momentum + m * (1.0f - momentum);
        moving_var[c] = moving_var[c] * momentum + unbiased * (1.0f - momentum);
    }
}

// y = act(gamma (x - mean) / sqrt(var + eps) + beta): batch statistics (training) or moving ones (inference)
__global__ void __launch_bounds__(kThreads) norm_apply_kernel(const float* __restrict__ x, int x_ld,
                                                              const float* __restrict__ mean,
                                                              const float* __restrict__ var,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps,
                                                              float* __restrict__ y, int y_ld, int64_t n_pix, int C,
                                                              int G, int act) {
    const int lane = threadIdx.x % G, row = threadIdx.x / G, rows = kThreads / G;
    float a[kMaxChunk], b[kMaxChunk];
#pragma unroll
    for (int k = 0; k < kMaxChunk; ++k) {
        const int c = lane + k * G;
        a[k] = 0.0f; b[k] = 0.0f;
        if (c < C) {
            a[k] = gamma[c] * rsqrtf(var[c] + eps);
            b[k] = beta[c] - mean[c] * a[k];
        }
    }
    for (int64_t p = (int64_t)blockIdx.x * rows + row; p < n_pix; p += (int64_t)gridDim.x * rows) {
#pragma unroll
        for (int k = 0; k < kMaxChunk; ++k) {
            const int c = lane + k * G;
            if (c < C) y[p * y_ld + c] = apply_act(fmaf(__ldg(x + p * x_ld + c), a[k], b[k]), act);
        }
    }
}

// sums[c] += sum_p dz, sums[C + c] += sum_p dz * xhat, dz = dy * act'(y)
__global__ void __launch_bounds__(kThreads) bn_bwd_reduce_kernel(const float* __restrict__ dy, int dy_ld,
                                                                 const float* __restrict__ x, int x_ld,
                                                                 const float* __restrict__ y, int y_ld,
                                                                 const float* __restrict__ mean,
                                                                 const float* __restrict__ var, float eps,
                                                                 int64_t n_pix, int C, int G, int act,
                                                                 float* __restrict__ sums) {
    __shared__ float sh[2 * 256];
    const int lane = threadIdx.x % G, row = threadIdx.x / G, rows = kThreads / G;
    float acc[2][kMaxChunk], mu[kMaxChunk], rs[kMaxChunk];
#pragma unroll
    for (int k = 0; k < kMaxChunk; ++k) {
        const int c = lane + k * G;
        acc[0][k] = 0.0f; acc[1][k] = 0.0f;
        mu[k] = c < C ? mean[c] : 0.0f;
        rs[k] = c < C ? rsqrtf(var[c] + eps) : 0.0f;
    }
    for (int64_t p = (int64_t)blockIdx.x * rows + row; p < n_pix; p += (int64_t)gridDim.x * rows) {
#pragma unroll
        for (int k = 0; k < kMaxChunk; ++k) {
            const int c = lane + k * G;
            if (c < C) {
                float dz = __ldg(dy + p * dy_ld + c);
                if (act != DL4DS_ACT_NONE) dz *= act_grad_from_out(__ldg(y + p * y_ld + c), act);
                acc[0][k] += dz;
                acc[1][k] += dz * (__ldg(x + p * x_ld + c) - mu[k]) * rs[k];
            }
        }
    }
    float* const dst[2] = {sums, sums + C};
    flush_channel_sums<2>(acc, lane, G, C, sh, dst);
}

// dx = gamma rstd (dz - mean(dz) - xhat mean(dz xhat)); block 0 also adds the two sums to dgamma / dbeta
__global__ void __launch_bounds__(kThreads) bn_bwd_apply_kernel(const float* __restrict__ dy, int dy_ld,
                                                                const float* __restrict__ x, int x_ld,
                                                                const float* __restrict__ y, int y_ld,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ var,
                                                                const float* __restrict__ gamma, float eps,
                                                                float* __restrict__ dx, int dx_ld, int64_t n_pix,
                                                                int C, int G, int act,
                                                                const float* __restrict__ sums,
                                                                float* __restrict__ dgamma,
                                                                float* __restrict__ dbeta) {
    const int lane = threadIdx.x % G, row = threadIdx.x / G, rows = kThreads / G;
    const float invM = 1.0f / (float)n_pix;
    float mu[kMaxChunk], rs[kMaxChunk], ga[kMaxChunk], m1[kMaxChunk], m2[kMaxChunk];
#pragma unroll
    for (int k = 0; k < kMaxChunk; ++k) {
        const int c = lane + k * G;
        const bool ok = c < C;
        mu[k] = ok ? mean[c] : 0.0f;
        rs[k] = ok ? rsqrtf(var[c] + eps) : 0.0f;
        ga[k] = ok ? gamma[c] * rs[k] : 0.0f;
        m1[k] = ok ? sums[c] * invM : 0.0f;
        m2[k] = ok ? sums[C + c] * invM : 0.0f;
    }
    if (blockIdx.x == 0 && dgamma) {
        for (int c = threadIdx.x; c < C; c += kThreads) {
            dbeta[c] += sums[c];
            dgamma[c] += sums[C + c];
        }
    }
    for (int64_t p = (int64_t)blockIdx.x * rows + row; p < n_pix; p += (int64_t)gridDim.x * rows) {
#pragma unroll
        for (int k = 0; k < kMaxChunk; ++k) {
            const int c = lane + k * G;
            if (c < C) {
                float dz = __ldg(dy + p * dy_ld + c);
                if (act != DL4DS_ACT_NONE) dz *= act_grad_from_out(__ldg(y + p * y_ld + c), act);
                const float xh = (__ldg(x + p * x_ld + c) - mu[k]) * rs[k];
                dx[p * dx_ld + c] = ga[k] * (dz - m1[k] - xh * m2[k]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// layer norm over the channel axis (one pixel = one normalisation group)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) ln_fwd_kernel(const float* __restrict__ x, int x_ld,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float eps,
                                                          float* __restrict__ y, int y_ld, int64_t n_pix, int C,
                                                          int G, int act) {
    const int lane = threadIdx.x % G, row = threadIdx.x / G, rows = kThreads / G;
    const float invC = 1.0f / (float)C;
    float ga[kMaxChunk], be[kMaxChunk];
#pragma unroll
    for (int k = 0; k < kMaxChunk; ++k) {
        const int c = lane + k * G;
        ga[k] = c < C ? gamma[c] : 0.0f;
        be[k] = c < C ? beta[c] : 0.0f;
    }
    for (int64_t base = (int64_t)blockIdx.x * rows; base < n_pix; base += (int64_t)gridDim.x * rows) {
        const int64_t p = base + row;
        const bool live = p < n_pix;
        float v[kMaxChunk], s = 0.0f;
#pragma unroll
        for (int k = 0; k < kMaxChunk; ++k) {
            const int c = lane + k * G;
            v[k] = (live && c < C) ? __ldg(x + p * x_ld + c) : 0.0f;
            s += v[k];
        }
        const float mu = group_sum(s, G) * invC;
        float ss = 0.0f;
#pragma unroll
        for (int k = 0; k < kMaxChunk; ++k) {
            const int c = lane + k * G;
            const float d = c < C ? v[k] - mu : 0.0f;
            ss = fmaf(d, d, ss);
        }
        const float rstd = rsqrtf(group_sum(ss, G) * invC + eps);
        if (live) {
#pragma unroll
            for (int k = 0; k < kMaxChunk; ++k) {
                const int c = lane + k * G;
                if (c < C) y[p * y_ld + c] = apply_act(fmaf((v[k] - mu) * rstd, ga[k], be[k]), act);
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads) ln_bwd_kernel(const float* __restrict__ dy, int dy_ld,
                                                          const float* __restrict__ x, int x_ld,
                                                          const float* __restrict__ y, int y_ld,
                                                          const float* __restrict__ gamma, float eps,
                                                          float* __restrict__ dx, int dx_ld,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                          int64_t n_pix, int C, int G, int act) {
    __shared__ float sh[2 * 256];
    const int lane = threadIdx.x % G, row = threadIdx.x / G, rows = kThreads / G;
    const float invC = 1.0f / (float)C;
    float ga[kMaxChunk], acc[2][kMaxChunk];
#pragma unroll
    for (int k = 0; k < kMaxChunk; ++k) {
        const int c = lane + k * G;
        ga[k] = c < C ? gamma[c] : 0.0f;
        acc[0][k] = 0.0f; acc[1][k] = 0.0f;
    }
    for (int64_t base = (int64_t)blockIdx.x * rows; base < n_pix; base += (int64_t)gridDim.x * rows) {
        const int64_t p = base + row;
        const bool live = p < n_pix;
        float v[kMaxChunk], dz[kMaxChunk], s = 0.0f;
#pragma unroll
        for (int k = 0; k < kMaxChunk; ++k) {
            const int c = lane + k * G;
            const bool ok = live && c < C;
            v[k] = ok ? __ldg(x + p * x_ld + c) : 0.0f;
            dz[k] = ok ? __ldg(dy + p * dy_ld + c) : 0.0f;
            if (ok && act != DL4DS_ACT_NONE) dz[k] *= act_grad_from_out(__ldg(y + p * y_ld + c), act);
            s += v[k];
        }
        const float mu = group_sum(s, G) * invC;
        float ss = 0.0f;
#pragma unroll
        for (int k = 0; k < kMaxChunk; ++k) {
            const int c = lane + k * G;
            const float d = c < C ? v[k] - mu : 0.0f;
            ss = fmaf(d, d, ss);
        }
        const float rstd = rsqrtf(group_sum(ss, G) * invC + eps);
        float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
        for (int k = 0; k < kMaxChunk; ++k) {
            const int c = lane + k * G;
            const float xh = c < C ? (v[k] - mu) * rstd : 0.0f;
            v[k] = xh;
            acc[0][k] += dz[k];             // d beta
            acc[1][k] += dz[k] * xh;        // d gamma
            const float dxh = dz[k] * ga[k];
            s1 += dxh;
            s2 = fmaf(dxh, xh, s2);
        }
        const float m1 = group_sum(s1, G) * invC, m2 = group_sum(s2, G) * invC;
        if (live && dx) {
#pragma unroll
            for (int k = 0; k < kMaxChunk; ++k) {
                const int c = lane + k * G;
                if (c < C) dx[p * dx_ld + c] = rstd * (dz[k] * ga[k] - m1 - v[k] * m2);
            }
        }
    }
    if (dgamma) {
        float* const dst[2] = {dbeta, dgamma};
        flush_channel_sums<2>(acc, lane, G, C, sh, dst);
    }
}

// ---------------------------------------------------------------------------------------------------------
// float4 variants (C % 4 == 0, pitches % 4 == 0, 16-byte aligned pointers): one lane = four consecutive channels.
// Batch norm: Q = C/4 lanes per pixel, 256/Q pixels per block step (per-channel constants in registers).
// Layer norm: G = next power of two >= Q lanes per pixel (C <= 128) so the per-pixel sums are xor-shuffles.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 act4(float4 v, int act) {
    return make_float4(apply_act(v.x, act), apply_act(v.y, act), apply_act(v.z, act), apply_act(v.w, act));
}
__device__ __forceinline__ float4 actgrad4(float4 dy, float4 y, int act) {
    return make_float4(dy.x * act_grad_from_out(y.x, act), dy.y * act_grad_from_out(y.y, act),
                       dy.z * act_grad_from_out(y.z, act), dy.w * act_grad_from_out(y.w, act));
}

template <int NACC>
__device__ __forceinline__ void flush_quad_sums(const float4 (&acc)[NACC], int q, bool on, int C, float* sh,
                                                float* const (&dst)[NACC]) {
    for (int i = threadIdx.x; i < NACC * 256; i += kThreads) sh[i] = 0.0f;
    __syncthreads();
    if (on) {
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
            atomicAdd(sh + a * 256 + 4 * q + 0, acc[a].x);
            atomicAdd(sh + a * 256 + 4 * q + 1, acc[a].y);
            atomicAdd(sh + a * 256 + 4 * q + 2, acc[a].z);
            atomicAdd(sh + a * 256 + 4 * q + 3, acc[a].w);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NACC * 256; i += kThreads) {
        const int a = i / 256, c = i % 256;
        if (c < C) atomicAdd(dst[a] + c, sh[i]);
    }
}

__global__ void __launch_bounds__(kThreads) bn_stats_v4_kernel(const float* __restrict__ x, int x_ld, int64_t n_pix,
                                                               int C, float* __restrict__ sums, int pass) {
    __shared__ float sh[256];
    const int Q = C >> 2, rows = kThreads / Q, q = threadIdx.x % Q, row = threadIdx.x / Q;
    const bool on = row < rows;
    float4 acc[1] = {make_float4(0.f, 0.f, 0.f, 0.f)};
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pass == 1) {
        const float inv = 1.0f / (float)n_pix;
        mu = make_float4(sums[4 * q] * inv, sums[4 * q + 1] * inv, sums[4 * q + 2] * inv, sums[4 * q + 3] * inv);
    }
    if (on) {
        for (int64_t p = (int64_t)blockIdx.x * rows + row; p < n_pix; p += (int64_t)gridDim.x * rows) {
            const float4 v = ld4(x + p * x_ld + 4 * q);
            const float dx = v.x - mu.x, dy = v.y - mu.y, dz = v.z - mu.z, dw = v.w - mu.w;
            if (pass == 0) { acc[0].x += dx; acc[0].y += dy; acc[0].z += dz; acc[0].w += dw; }
            else { acc[0].x = fmaf(dx, dx, acc[0].x); acc[0].y = fmaf(dy, dy, acc[0].y);
                   acc[0].z = fmaf(dz, dz, acc[0].z); acc[0].w = fmaf(dw, dw, acc[0].w); }
        }
    }
    float* const dst[1] = {sums + (pass == 0 ? 0 : C)};
    flush_quad_sums<1>(acc, q, on, C, sh, dst);
}

__global__ void __launch_bounds__(kThreads) norm_apply_v4_kernel(const float* __restrict__ x, int x_ld,
                                                                 const float* __restrict__ mean,
                                                                 const float* __restrict__ var,
                                                                 const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float eps,
                                                                 float* __restrict__ y, int y_ld, int64_t n_pix,
                                                                 int C, int act) {
    const int Q = C >> 2, rows = kThreads / Q, q = threadIdx.x % Q, row = threadIdx.x / Q;
    if (row >= rows) return;
    float a[4], b[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = 4 * q + k;
        a[k] = gamma[c] * rsqrtf(var[c] + eps);
        b[k] = beta[c] - mean[c] * a[k];
    }
    for (int64_t p = (int64_t)blockIdx.x * rows + row; p < n_pix; p += (int64_t)gridDim.x * rows) {
        const float4 v = ld4(x + p * x_ld + 4 * q);
        st4(y + p * y_ld + 4 * q, act4(make_float4(fmaf(v.x, a[0], b[0]), fmaf(v.y, a[1], b[1]), fmaf(v.z, a[2], b[2]),
                                                   fmaf(v.w, a[3], b[3])), act));
    }
}

__global__ void __launch_bounds__(kThreads) bn_bwd_reduce_v4_kernel(const float* __restrict__ dy, int dy_ld,
                                                                    const float* __restrict__ x, int x_ld,
                                                                    const float* __restrict__ y, int y_ld,
                                                                    const float* __restrict__ mean,
                                                                    const float* __restrict__ var, float eps,
                                                                    int64_t n_pix, int C, int act,
                                                                    float* __restrict__ sums) {
    __shared__ float sh[2 * 256];
    const int Q = C >> 2, rows = kThreads / Q, q = threadIdx.x % Q, row = threadIdx.x / Q;
    const bool on = row < rows;
    float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
    const float4 mu = ld4(mean + 4 * q);
    const float4 vr = ld4(var + 4 * q);
    const float4 rs = make_float4(rsqrtf(vr.x + eps), rsqrtf(vr.y + eps), rsqrtf(vr.z + eps), rsqrtf(vr.w + eps));
    if (on) {
        for (int64_t p = (int64_t)blockIdx.x * rows + row; p < n_pix; p += (int64_t)gridDim.x * rows) {
            float4 dz = ld4(dy + p * dy_ld + 4 * q);
            if (act != DL4DS_ACT_NONE) dz = actgrad4(dz, ld4(y + p * y_ld + 4 * q), act);
            const float4 v = ld4(x + p * x_ld + 4 * q);
            acc[0].x += dz.x; acc[0].y += dz.y; acc[0].z += dz.z; acc[0].w += dz.w;
            acc[1].x = fmaf(dz.x, (v.x - mu.x) * rs.x, acc[1].x);
            acc[1].y = fmaf(dz.y, (v.y - mu.y) * rs.y, acc[1].y);
            acc[1].z = fmaf(dz.z, (v.z - mu.z) * rs.z, acc[1].z);
            acc[1].w = fmaf(dz.w, (v.w - mu.w) * rs.w, acc[1].w);
        }
    }
    float* const dst[2] = {sums, sums + C};
    flush_quad_sums<2>(acc, q, on, C, sh, dst);
}

__global__ void __launch_bounds__(kThreads) bn_bwd_apply_v4_kernel(const float* __restrict__ dy, int dy_ld,
                                                                   const float* __restrict__ x, int x_ld,
                                                                   const float* __restrict__ y, int y_ld,
                                                                   const float* __restrict__ mean,
                                                                   const float* __restrict__ var,
                                                                   const float* __restrict__ gamma, float eps,
                                                                   float* __restrict__ dx, int dx_ld, int64_t n_pix,
                                                                   int C, int act, const float* __restrict__ sums,
                                                                   float* __restrict__ dgamma,
                                                                   float* __restrict__ dbeta) {
    const int Q = C >> 2, rows = kThreads / Q, q = threadIdx.x % Q, row = threadIdx.x / Q;
    if (blockIdx.x == 0 && dgamma) {
        for (int c = threadIdx.x; c < C; c += kThreads) {
            dbeta[c] += sums[c];
            dgamma[c] += sums[C + c];
        }
    }
    if (row >= rows) return;
    const float invM = 1.0f / (float)n_pix;
    float mu[4], rs[4], ga[4], m1[4], m2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = 4 * q + k;
        mu[k] = mean[c];
        rs[k] = rsqrtf(var[c] + eps);
        ga[k] = gamma[c] * rs[k];
        m1[k] = sums[c] * invM;
        m2[k] = sums[C + c] * invM;
    }
    for (int64_t p = (int64_t)blockIdx.x * rows + row; p < n_pix; p += (int64_t)gridDim.x * rows) {
        float4 dz = ld4(dy + p * dy_ld + 4 * q);
        if (act != DL4DS_ACT_NONE) dz = actgrad4(dz, ld4(y + p * y_ld + 4 * q), act);
        const float4 v = ld4(x + p * x_ld + 4 * q);
        st4(dx + p * dx_ld + 4 * q,
            make_float4(ga[0] * (dz.x - m1[0] - (v.x - mu[0]) * rs[0] * m2[0]),
                        ga[1] * (dz.y - m1[1] - (v.y - mu[1]) * rs[1] * m2[1]),
                        ga[2] * (dz.z - m1[2] - (v.z - mu[2]) * rs[2] * m2[2]),
                        ga[3] * (dz.w - m1[3] - (v.w - mu[3]) * rs[3] * m2[3])));
    }
}

__global__ void __launch_bounds__(kThreads) ln_fwd_v4_kernel(const float* __restrict__ x, int x_ld,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps,
                                                             float* __restrict__ y, int y_ld, int64_t n_pix, int C,
                                                             int G, int act) {
    const int lane = threadIdx.x % G, row = threadIdx.x / G, rows = kThreads / G;
    const bool has = 4 * lane < C;
    const float invC = 1.0f / (float)C;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 ga = has ? ld4(gamma + 4 * lane) : z, be = has ? ld4(beta + 4 * lane) : z;
    for (int64_t base = (int64_t)blockIdx.x * rows; base < n_pix; base += (int64_t)gridDim.x * rows) {
        const int64_t p = base + row;
        const bool live = has && p < n_pix;
        const float4 v = live ? ld4(x + p * x_ld + 4 * lane) : z;
        const float mu = group_sum(v.x + v.y + v.z + v.w, G) * invC;
        const float4 d = has ? make_float4(v.x - mu, v.y - mu, v.z - mu, v.w - mu) : z;
        const float rstd = rsqrtf(group_sum(d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w, G) * invC + eps);
        if (live)
            st4(y + p * y_ld + 4 * lane, act4(make_float4(fmaf(d.x * rstd, ga.x, be.x), fmaf(d.y * rstd, ga.y, be.y),
                                                          fmaf(d.z * rstd, ga.z, be.z), fmaf(d.w * rstd, ga.w, be.w)),
                                              act));
    }
}

__global__ void __launch_bounds__(kThreads) ln_bwd_v4_kernel(const float* __restrict__ dy, int dy_ld,
                                                             const float* __restrict__ x, int x_ld,
                                                             const float* __restrict__ y, int y_ld,
                                                             const float* __restrict__ gamma, float eps,
                                                             float* __restrict__ dx, int dx_ld,
                                                             float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                             int64_t n_pix, int C, int G, int act) {
    __shared__ float sh[2 * 256];
    const int lane = threadIdx.x % G, row = threadIdx.x / G, rows = kThreads / G;
    const bool has = 4 * lane < C;
    const float invC = 1.0f / (float)C;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 ga = has ? ld4(gamma + 4 * lane) : z;
    float4 acc[2] = {z, z};
    for (int64_t base = (int64_t)blockIdx.x * rows; base < n_pix; base += (int64_t)gridDim.x * rows) {
        const int64_t p = base + row;
        const bool live = has && p < n_pix;
        const float4 v = live ? ld4(x + p * x_ld + 4 * lane) : z;
        float4 dz = live ? ld4(dy + p * dy_ld + 4 * lane) : z;
        if (live && act != DL4DS_ACT_NONE) dz = actgrad4(dz, ld4(y + p * y_ld + 4 * lane), act);
        const float mu = group_sum(v.x + v.y + v.z + v.w, G) * invC;
        const float4 d = has ? make_float4(v.x - mu, v.y - mu, v.z - mu, v.w - mu) : z;
        const float rstd = rsqrtf(group_sum(d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w, G) * invC + eps);
        const float4 xh = make_float4(d.x * rstd, d.y * rstd, d.z * rstd, d.w * rstd);
        acc[0].x += dz.x; acc[0].y += dz.y; acc[0].z += dz.z; acc[0].w += dz.w;
        acc[1].x = fmaf(dz.x, xh.x, acc[1].x); acc[1].y = fmaf(dz.y, xh.y, acc[1].y);
        acc[1].z = fmaf(dz.z, xh.z, acc[1].z); acc[1].w = fmaf(dz.w, xh.w, acc[1].w);
        const float4 g = make_float4(dz.x * ga.x, dz.y * ga.y, dz.z * ga.z, dz.w * ga.w);
        const float m1 = group_sum(g.x + g.y + g.z + g.w, G) * invC;
        const float m2 = group_sum(g.x * xh.x + g.y * xh.y + g.z * xh.z + g.w * xh.w, G) * invC;
        if (live && dx)
            st4(dx + p * dx_ld + 4 * lane, make_float4(rstd * (g.x - m1 - xh.x * m2), rstd * (g.y - m1 - xh.y * m2),
                                                       rstd * (g.z - m1 - xh.z * m2), rstd * (g.w - m1 - xh.w * m2)));
    }
    if (dgamma) {
        float* const dst[2] = {dbeta, dgamma};
        flush_quad_sums<2>(acc, lane, has, C, sh, dst);
    }
}

bool vec4_ok(int C, std::initializer_list<int> lds, std::initializer_list<const void*> ptrs) {
    if (C % 4) return false;
    for (int ld : lds)
        if (ld % 4) return false;
    for (const void* p : ptrs)
        if (p && (reinterpret_cast<uintptr_t>(p) & 15)) return false;
    return true;
}

int grid_quads(int64_t n_pix, int lanes_per_pixel, int max_blocks) {
    const int rows = kThreads / lanes_per_pixel;
    return (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(n_pix, (int64_t)rows * 2), max_blocks));
}

int ln_group(int C) {      // lanes per pixel for the float4 layer norm, 0 if it does not apply
    if (C % 4 || C > 128) return 0;
    int g = 1;
    while (4 * g < C) g <<= 1;
    return g;
}

int grid_rows(int64_t n_pix, int G) {
    const int rows = kThreads / G;
    const int64_t blocks = cdiv(n_pix, (int64_t)rows * 4);
    return (int)std::max<int64_t>(1, std::min<int64_t>(blocks, 8 * kNumSMs));
}

}  // namespace
}  // namespace dl4ds

using namespace dl4ds;

#define NORM_CHECK(what)                                                                                       \
    DL4DS_REQUIRE(n_pix > 0 && C > 0, DL4DS_E_SHAPE, what ": n_pix <= 0 or C <= 0");                           \
    DL4DS_REQUIRE(C <= 256, DL4DS_E_UNSUPPORTED, what ": C = %d > 256 channels", C);                           \
    DL4DS_REQUIRE(act >= DL4DS_ACT_NONE && act <= DL4DS_ACT_TANH, DL4DS_E_BADARG, what ": bad activation")

extern "C" {

int dl4ds_batchnorm_stats(const float* x, int x_ld, int64_t n_pix, int C, float* mean, float* var,
                          float* moving_mean, float* moving_var, float momentum, float* ws, void* stream) {
    const int act = 0;
    NORM_CHECK("batchnorm_stats");
    DL4DS_REQUIRE(x && mean && var && ws, DL4DS_E_BADARG, "batchnorm_stats: null pointer");
    DL4DS_REQUIRE((moving_mean == nullptr) == (moving_var == nullptr), DL4DS_E_BADARG,
                  "batchnorm_stats: moving_mean / moving_var must come together");
    cudaStream_t st = as_stream(stream);
    const int G = group_width(C), grid = grid_rows(n_pix, G);
    cudaMemsetAsync(ws, 0, 2 * (size_t)C * sizeof(float), st);
    if (vec4_ok(C, {x_ld}, {x})) {
        const int g4 = grid_quads(n_pix, C / 4, 4 * kNumSMs);
        bn_stats_v4_kernel<<<g4, kThreads, 0, st>>>(x, x_ld, n_pix, C, ws, 0);
        bn_stats_v4_kernel<<<g4, kThreads, 0, st>>>(x, x_ld, n_pix, C, ws, 1);
    } else {
        bn_stats_kernel<<<grid, kThreads, 0, st>>>(x, x_ld, n_pix, C, G, ws, 0);
        bn_stats_kernel<<<grid, kThreads, 0, st>>>(x, x_ld, n_pix, C, G, ws, 1);
    }
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(ws, C, (float)n_pix, mean, var, moving_mean, moving_var,
                                                        momentum);
    return check_launch("batchnorm_stats");
}

int dl4ds_norm_apply(const float* x, int x_ld, const float* mean, const float* var, const float* gamma,
                     const float* beta, float eps, float* y, int y_ld, int64_t n_pix, int C, int act,
                     void* stream) {
    NORM_CHECK("norm_apply");
    DL4DS_REQUIRE(x && mean && var && gamma && beta && y, DL4DS_E_BADARG, "norm_apply: null pointer");
    const int G = group_width(C);
    if (vec4_ok(C, {x_ld, y_ld}, {x, y}))
        norm_apply_v4_kernel<<<grid_quads(n_pix, C / 4, 16 * kNumSMs), kThreads, 0, as_stream(stream)>>>(
            x, x_ld, mean, var, gamma, beta, eps, y, y_ld, n_pix, C, act);
    else
        norm_apply_kernel<<<grid_rows(n_pix, G), kThreads, 0, as_stream(stream)>>>(x, x_ld, mean, var, gamma, beta,
                                                                                    eps, y, y_ld, n_pix, C, G, act);
    return check_launch("norm_apply");
}

int dl4ds_batchnorm_bwd(const float* dy, int dy_ld, const float* x, int x_ld, const float* y, int y_ld,
                        const float* mean, const float* var, const float* gamma, float eps, float* dx, int dx_ld,
                        float* dgamma, float* dbeta, float* ws, int64_t n_pix, int C, int act, void* stream) {
    NORM_CHECK("batchnorm_bwd");
    DL4DS_REQUIRE(dy && x && mean && var && gamma && dx && ws, DL4DS_E_BADARG, "batchnorm_bwd: null pointer");
    DL4DS_REQUIRE(act == DL4DS_ACT_NONE || y, DL4DS_E_BADARG, "batchnorm_bwd: y is needed for the activation");
    DL4DS_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), DL4DS_E_BADARG,
                  "batchnorm_bwd: dgamma / dbeta must come together");
    cudaStream_t st = as_stream(stream);
    const int G = group_width(C), grid = grid_rows(n_pix, G);
    cudaMemsetAsync(ws, 0, 2 * (size_t)C * sizeof(float), st);
    if (vec4_ok(C, {dy_ld, x_ld, y_ld, dx_ld}, {dy, x, y, dx, mean, var})) {
        bn_bwd_reduce_v4_kernel<<<grid_quads(n_pix, C / 4, 4 * kNumSMs), kThreads, 0, st>>>(
            dy, dy_ld, x, x_ld, y, y_ld, mean, var, eps, n_pix, C, act, ws);
        bn_bwd_apply_v4_kernel<<<grid_quads(n_pix, C / 4, 16 * kNumSMs), kThreads, 0, st>>>(
            dy, dy_ld, x, x_ld, y, y_ld, mean, var, gamma, eps, dx, dx_ld, n_pix, C, act, ws, dgamma, dbeta);
        return check_launch("batchnorm_bwd");
    }
    bn_bwd_reduce_kernel<<<grid, kThreads, 0, st>>>(dy, dy_ld, x, x_ld, y, y_ld, mean, var, eps, n_pix, C, G, act, ws);
    bn_bwd_apply_kernel<<<grid, kThreads, 0, st>>>(dy, dy_ld, x, x_ld, y, y_ld, mean, var, gamma, eps, dx, dx_ld,
                                                   n_pix, C, G, act, ws, dgamma, dbeta);
    return check_launch("batchnorm_bwd");
}

int dl4ds_layernorm_fwd(const float* x, int x_ld, const float* gamma, const float* beta, float eps, float* y,
                        int y_ld, int64_t n_pix, int C, int act, void* stream) {
    NORM_CHECK("layernorm_fwd");
    DL4DS_REQUIRE(x && gamma && beta && y, DL4DS_E_BADARG, "layernorm_fwd: null pointer");
    const int G = group_width(C);
    const int g4 = ln_group(C);
    if (g4 && vec4_ok(C, {x_ld, y_ld}, {x, y, gamma, beta}))
        ln_fwd_v4_kernel<<<grid_quads(n_pix, g4, 16 * kNumSMs), kThreads, 0, as_stream(stream)>>>(
            x, x_ld, gamma, beta, eps, y, y_ld, n_pix, C, g4, act);
    else
        ln_fwd_kernel<<<grid_rows(n_pix, G), kThreads, 0, as_stream(stream)>>>(x, x_ld, gamma, beta, eps, y, y_ld,
                                                                                n_pix, C, G, act);
    return check_launch("layernorm_fwd");
}

int dl4ds_layernorm_bwd(const float* dy, int dy_ld, const float* x, int x_ld, const float* y, int y_ld,
                        const float* gamma, float eps, float* dx, int dx_ld, float* dgamma, float* dbeta,
                        int64_t n_pix, int C, int act, void* stream) {
    NORM_CHECK("layernorm_bwd");
    DL4DS_REQUIRE(dy && x && gamma, DL4DS_E_BADARG, "layernorm_bwd: null pointer");
    DL4DS_REQUIRE(act == DL4DS_ACT_NONE || y, DL4DS_E_BADARG, "layernorm_bwd: y is needed for the activation");
    DL4DS_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), DL4DS_E_BADARG,
                  "layernorm_bwd: dgamma / dbeta must come together");
    const int G = group_width(C);
    const int g4 = ln_group(C);
    if (g4 && vec4_ok(C, {dy_ld, x_ld, y_ld, dx_ld}, {dy, x, y, dx, gamma}))
        ln_bwd_v4_kernel<<<grid_quads(n_pix, g4, 4 * kNumSMs), kThreads, 0, as_stream(stream)>>>(
            dy, dy_ld, x, x_ld, y, y_ld, gamma, eps, dx, dx_ld, dgamma, dbeta, n_pix, C, g4, act);
    else
        ln_bwd_kernel<<<grid_rows(n_pix, G), kThreads, 0, as_stream(stream)>>>(dy, dy_ld, x, x_ld, y, y_ld, gamma, eps,
                                                                                dx, dx_ld, dgamma, dbeta, n_pix, C, G,
                                                                                act);
    return check_launch("layernorm_bwd");
}

}  // extern "C"
