// HBM-bound graph ops of the DL4DS hot path (everything that is not a convolution GEMM):
// epilogue backward (bias/activation/space_to_depth), Add / Concatenate slices, ChannelAttention2D,
// pixel + adversarial losses, TF-Adam, block-mean coarsening, bilinear resize, 2x2 max-pool,
// LocallyConnected2D 1x1, ConvLSTM gate math, global average pooling.
// All tensors fp32 NHWC with an explicit channel pitch (`*_ld`).  Reductions use warp shuffles, a
// shared-memory stage and one atomic per CTA per output (blocks.py / losses.py call sites are named
// at each entry point in include/dl4ds_b200.h).
#include "common.cuh"

namespace dl4ds {

static inline int grid_for(int64_t work_items, int per_block, int max_blocks = 8 * kNumSMs) {
    int64_t g = cdiv(work_items, per_block);
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (int)g;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// -------------------------------------------------------------------------------------------------
// bias + activation (+ depth_to_space) backward
// -------------------------------------------------------------------------------------------------
// block = (TX, 256/TX); thread (tx,ty) owns channels c = tx + k*TX (k < KS) and pixels ty, ty+PY, ...
template <int KS>
__global__ void __launch_bounds__(256) bias_act_bwd_kernel(
    const float* __restrict__ dy, int dy_ld, const float* __restrict__ y, int y_ld,
    float* __restrict__ dz, int dz_ld, float* __restrict__ dbias,
    int64_t n_pix, int Ho, int Wo, int C, int act, int r) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    __shared__ float red[256];
    const int TX = blockDim.x, PY = blockDim.y;
    const int tx = threadIdx.x, ty = threadIdx.y;
    float acc[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) acc[k] = 0.0f;
    const int Cd = (r > 1) ? C / (r * r) : C;
    for (int64_t p = (int64_t)blockIdx.x * PY + ty; p < n_pix; p += (int64_t)gridDim.x * PY) {
        int64_t src_base = p;   // pixel index into dy
        int n = 0, oy = 0, ox = 0;
        if (r > 1) {
            const int hw = Ho * Wo;
            n = (int)(p / hw);
            const int rem = (int)(p - (int64_t)n * hw);
            oy = rem / Wo; ox = rem - oy * Wo;
        }
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            const int c = tx + k * TX;
            if (c < C) {
                float g;
                if (r > 1) {
                    const int grp = c / Cd, cc = c - grp * Cd;
                    const int di = grp / r, dj = grp - di * r;
                    const int64_t hp = ((int64_t)(n * Ho * r + oy * r + di)) * (Wo * r) + ox * r + dj;
                    g = __ldg(dy + hp * dy_ld + cc);
                    if (act != DL4DS_ACT_NONE) g *= act_grad_from_out(__ldg(y + hp * y_ld + cc), act);
                } else {
                    g = __ldg(dy + src_base * dy_ld + c);
                    if (act != DL4DS_ACT_NONE) g *= act_grad_from_out(__ldg(y + p * y_ld + c), act);
                }
                if (dz) dz[p * dz_ld + c] = g;
                acc[k] += g;
            }
        }
    }
    if (dbias == nullptr) return;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        red[ty * TX + tx] = acc[k];
        __syncthreads();
        if (ty == 0) {
            float s = 0.0f;
            for (int j = 0; j < PY; ++j) s += red[j * TX + tx];
            const int c = tx + k * TX;
            if (c < C) atomicAdd(dbias + c, s);
        }
        __syncthreads();
    }
}

// -------------------------------------------------------------------------------------------------
// generic element-wise kernels over (n_pix, C) with pitches
// -------------------------------------------------------------------------------------------------
__global__ void add_kernel(const float* __restrict__ a, int a_ld, const float* __restrict__ b, int b_ld,
                           float* __restrict__ out, int out_ld, int64_t n_pix, int C, int act) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int64_t total = n_pix * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / C;
        const int c = (int)(i - p * C);
        out[p * out_ld + c] = apply_act(a[p * a_ld + c] + b[p * b_ld + c], act);
    }
}

__global__ void add_vec4_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                float4* __restrict__ out, int64_t n4, int act) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float4 u = __ldg(a + i), v = __ldg(b + i);
        float4 o;
        o.x = apply_act(u.x + v.x, act); o.y = apply_act(u.y + v.y, act);
        o.z = apply_act(u.z + v.z, act); o.w = apply_act(u.w + v.w, act);
        out[i] = o;
    }
}

__global__ void copy_channels_kernel(const float* __restrict__ src, int src_ld, float* __restrict__ dst,
                                     int dst_ld, int64_t n_pix, int C, int accumulate) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int64_t total = n_pix * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / C;
        const int c = (int)(i - p * C);
        const float v = __ldg(src + p * src_ld + c);
        float* d = dst + p * dst_ld + c;
        *d = accumulate ? (*d + v) : v;
    }
}

__global__ void copy_channels_vec4_kernel(const float* __restrict__ src, int src_ld,
                                          float* __restrict__ dst, int dst_ld, int64_t n_pix, int C4,
                                          int accumulate) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int64_t total = n_pix * C4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / C4;
        const int c = (int)(i - p * C4) * 4;
        float4 v = __ldg(reinterpret_cast<const float4*>(src + p * src_ld + c));
        float4* d = reinterpret_cast<float4*>(dst + p * dst_ld + c);
        if (accumulate) {
            const float4 o = *d;
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        *d = v;
    }
}

__global__ void act_fwd_kernel(const float* __restrict__ x, int x_ld, float* __restrict__ y, int y_ld,
                               int64_t n_pix, int C, int act) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int64_t total = n_pix * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / C;
        const int c = (int)(i - p * C);
        y[p * y_ld + c] = apply_act(__ldg(x + p * x_ld + c), act);
    }
}

__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b,
                           float* __restrict__ out, int64_t n) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = a[i] * b[i];
}

__global__ void axpby_kernel(float a, const float* __restrict__ x, float b, float* __restrict__ y,
                             int64_t n) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        y[i] = a * x[i] + (b == 0.0f ? 0.0f : b * y[i]);
}

// -------------------------------------------------------------------------------------------------
// grouped pixel reductions: out[g,c] (+)= sum_{p in group g} f(p,c)
// pixel p belongs to group (p / (ppg*inner))*inner + p % inner   (inner = 1 for NHWC images pooled
// over H*W; inner = W for the 5-D (T,H) pooling quirk of blocks.py:587 on NTHWC tensors)
// block = (TX, 256/TX) as in bias_act_bwd; grid = (chunks, n_groups)
// -------------------------------------------------------------------------------------------------
template <int KS, bool kMulB>
__global__ void __launch_bounds__(256) group_sum_kernel(
    const float* __restrict__ a, int a_ld, const float* __restrict__ b, int b_ld,
    float* __restrict__ out, int64_t ppg, int inner, int C) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    __shared__ float red[256];
    const int TX = blockDim.x, PY = blockDim.y;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int g = blockIdx.y;
    const int go = g / inner, gi = g - go * inner;
    float acc[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) acc[k] = 0.0f;
    for (int64_t q = (int64_t)blockIdx.x * PY + ty; q < ppg; q += (int64_t)gridDim.x * PY) {
        const int64_t p = ((int64_t)go * ppg + q) * inner + gi;
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            const int c = tx + k * TX;
            if (c < C) {
                float v = __ldg(a + p * a_ld + c);
                if (kMulB) v *= __ldg(b + p * b_ld + c);
                acc[k] += v;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        red[ty * TX + tx] = acc[k];
        __syncthreads();
        if (ty == 0) {
            float s = 0.0f;
            for (int j = 0; j < PY; ++j) s += red[j * TX + tx];
            const int c = tx + k * TX;
            if (c < C) atomicAdd(out + (int64_t)g * C + c, s);
        }
        __syncthreads();
    }
}

// y[p,c] = x[p,c] * s[g(p),c]            (mode 0)
// dx[p,c] = dy[p,c]*s[g,c] + dm[g,c]*inv (mode 1)
// dx[p,c] = dm[g,c]*inv                  (mode 2)
__global__ void group_scale_kernel(const float* __restrict__ x, int x_ld, float* __restrict__ y,
                                   int y_ld, const float* __restrict__ s, const float* __restrict__ dm,
                                   float inv, int64_t n_pix, int64_t ppg, int inner, int C, int mode) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int64_t total = n_pix * C;
    const int64_t span = ppg * inner;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / C;
        const int c = (int)(i - p * C);
        const int64_t g = (p / span) * inner + (p % inner);
        float v;
        if (mode == 0) v = __ldg(x + p * x_ld + c) * __ldg(s + g * C + c);
        else if (mode == 1) v = __ldg(x + p * x_ld + c) * __ldg(s + g * C + c) + __ldg(dm + g * C + c) * inv;
        else v = __ldg(dm + g * C + c) * inv;
        y[p * y_ld + c] = v;
    }
}

// 16-byte variant (C % 4 == 0, pitches % 4 == 0, 16-byte aligned, < 2^31 float4 items): 32-bit index math, one
// float4 per thread and iteration (the scalar kernel above spends its time in 64-bit divisions: 55 us for the
// 33.5 MB HR tensor of the headline model, 5x the HBM roofline)
__global__ void __launch_bounds__(256) group_scale_vec4_kernel(
    const float* __restrict__ x, int x_ld, float* __restrict__ y, int y_ld, const float* __restrict__ s,
    const float* __restrict__ dm, float inv, int n_items, int ppg, int inner, int C, int mode) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int G = C >> 2;
    const int span = ppg * inner;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n_items; i += gridDim.x * 256) {
        const int p = i / G, c = (i - p * G) << 2;
        const int g = inner == 1 ? p / ppg : (p / span) * inner + (p % inner);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mode != 2) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(x + (int64_t)p * x_ld + c));
            const float4 sc = __ldg(reinterpret_cast<const float4*>(s + (int64_t)g * C + c));
            v = make_float4(a.x * sc.x, a.y * sc.y, a.z * sc.z, a.w * sc.w);
        }
        if (mode != 0) {
            const float4 d = __ldg(reinterpret_cast<const float4*>(dm + (int64_t)g * C + c));
            v.x += d.x * inv; v.y += d.y * inv; v.z += d.z * inv; v.w += d.w * inv;
        }
        *reinterpret_cast<float4*>(y + (int64_t)p * y_ld + c) = v;
    }
}

// 16-byte variant of group_sum_kernel for C % 4 == 0, C <= 64, inner == 1: block = (C/4, 256/(C/4)), 4 pixels in flight
template <bool kMulB>
__global__ void __launch_bounds__(256) group_sum_vec4_kernel(
    const float* __restrict__ a, int a_ld, const float* __restrict__ b, int b_ld,
    float* __restrict__ out, int ppg, int C) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    __shared__ float4 red[256];
    const int TX = blockDim.x, PY = blockDim.y;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int g = blockIdx.y;
    const int64_t p0 = (int64_t)g * ppg;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int step = gridDim.x * PY;
    int q = blockIdx.x * PY + ty;
    for (; q + 3 * step < ppg; q += 4 * step) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(a + (p0 + q + u * step) * a_ld) + tx);
        if (kMulB) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(b + (p0 + q + u * step) * b_ld) + tx);
                v[u].x *= w.x; v[u].y *= w.y; v[u].z *= w.z; v[u].w *= w.w;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    for (; q < ppg; q += step) {
        float4 v = __ldg(reinterpret_cast<const float4*>(a + (p0 + q) * a_ld) + tx);
        if (kMulB) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(b + (p0 + q) * b_ld) + tx);
            v.x *= w.x; v.y *= w.y; v.z *= w.z; v.w *= w.w;
        }
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    red[ty * TX + tx] = acc;
    __syncthreads();
    for (int h = PY >> 1; h > 0; h >>= 1) {
        if (ty < h) {
            const float4 t = red[(ty + h) * TX + tx];
            float4& r = red[ty * TX + tx];
            r.x += t.x; r.y += t.y; r.z += t.z; r.w += t.w;
        }
        __syncthreads();
    }
    if (ty == 0) {
        const float4 r = red[tx];
        float* o = out + (int64_t)g * C + tx * 4;
        atomicAdd(o + 0, r.x); atomicAdd(o + 1, r.y); atomicAdd(o + 2, r.z); atomicAdd(o + 3, r.w);
    }
}

// squeeze-excite MLP of ChannelAttention2D (blocks.py:582-593): one warp per group.
__global__ void attention_mlp_fwd_kernel(const float* __restrict__ pooled, float inv,
                                         const float* __restrict__ w1, const float* __restrict__ b1,
                                         const float* __restrict__ w2, const float* __restrict__ b2,
                                         float* __restrict__ hidden, float* __restrict__ scale,
                                         int n_groups, int C, int Cr) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int g = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (g >= n_groups) return;
    for (int j = 0; j < Cr; ++j) {
        float s = 0.0f;
        for (int c = lane; c < C; c += 32) s += pooled[(int64_t)g * C + c] * inv * __ldg(w1 + c * Cr + j);
        s = warp_sum(s);
        if (lane == 0) hidden[(int64_t)g * Cr + j] = fmaxf(s + __ldg(b1 + j), 0.0f);
    }
    __syncwarp();
    for (int c = lane; c < C; c += 32) {
        float s = __ldg(b2 + c);
        for (int j = 0; j < Cr; ++j) s += hidden[(int64_t)g * Cr + j] * __ldg(w2 + j * C + c);
        scale[(int64_t)g * C + c] = 1.0f / (1.0f + expf(-s));
    }
}

// backward of the MLP; dsum holds sum_p dy*x on entry and d(mean) on exit.  One warp per group.
__global__ void attention_mlp_bwd_kernel(const float* __restrict__ pooled, float inv,
                                         const float* __restrict__ w1, const float* __restrict__ w2,
                                         const float* __restrict__ hidden, const float* __restrict__ scale,
                                         float* __restrict__ dsum, float* __restrict__ dw1,
                                         float* __restrict__ db1, float* __restrict__ dw2,
                                         float* __restrict__ db2, int n_groups, int C, int Cr) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    extern __shared__ float sm[];   // per warp: dsig[C] + dhid[Cr]
    const int wib = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int g = blockIdx.x * (blockDim.x / 32) + wib;
    if (g >= n_groups) return;
    float* dsig = sm + wib * (C + Cr);
    float* dhid = dsig + C;
    for (int c = lane; c < C; c += 32) {
        const float s = scale[(int64_t)g * C + c];
        const float d = dsum[(int64_t)g * C + c] * s * (1.0f - s);
        dsig[c] = d;
        atomicAdd(db2 + c, d);
    }
    __syncwarp();
    for (int j = 0; j < Cr; ++j) {
        const float h = hidden[(int64_t)g * Cr + j];
        float s = 0.0f;
        for (int c = lane; c < C; c += 32) {
            s += dsig[c] * __ldg(w2 + j * C + c);
            atomicAdd(dw2 + j * C + c, h * dsig[c]);
        }
        s = warp_sum(s);
        if (lane == 0) {
            const float d = h > 0.0f ? s : 0.0f;
            dhid[j] = d;
            atomicAdd(db1 + j, d);
        }
    }
    __syncwarp();
    for (int c = lane; c < C; c += 32) {
        const float mean = pooled[(int64_t)g * C + c] * inv;
        float s = 0.0f;
        for (int j = 0; j < Cr; ++j) {
            s += dhid[j] * __ldg(w1 + c * Cr + j);
            atomicAdd(dw1 + c * Cr + j, mean * dhid[j]);
        }
        dsum[(int64_t)g * C + c] = s;
    }
}

// -------------------------------------------------------------------------------------------------
// losses
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pixel_loss_kernel(const float* __restrict__ yp,
                                                         const float* __restrict__ yt,
                                                         float* __restrict__ loss_out,
                                                         float* __restrict__ dy, int64_t n, int kind,
                                                         float scale, int accumulate) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    __shared__ float red[8];
    const float invn = 1.0f / (float)n;
    float acc = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float d = __ldg(yp + i) - __ldg(yt + i);
        if (kind == 0) {
            acc += fabsf(d);
            if (dy) {
                const float g = (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f)) * (scale * invn);
                dy[i] = accumulate ? dy[i] + g : g;
            }
        } else {
            acc += d * d;
            if (dy) {
                const float g = 2.0f * d * (scale * invn);
                dy[i] = accumulate ? dy[i] + g : g;
            }
        }
    }
    acc = warp_sum(acc);
    if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < 8 ? red[threadIdx.x] : 0.0f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(loss_out, v * invn * scale);
    }
}

// Keras BinaryCrossentropy(from_logits=False): p = clip(p, eps, 1-eps);
// -mean(t*log(p+eps) + (1-t)*log(1-p+eps)), eps = 1e-7
__global__ void bce_loss_kernel(const float* __restrict__ p, float target, float* __restrict__ loss_out,
                                float* __restrict__ dp, int64_t n, float scale, int accumulate) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const float eps = 1e-7f;
    const float invn = 1.0f / (float)n;
    float acc = 0.0f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float raw = p[i];
        const float pc = fminf(fmaxf(raw, eps), 1.0f - eps);
        acc += -(target * logf(pc + eps) + (1.0f - target) * logf(1.0f - pc + eps));
        if (dp) {
            float g = 0.0f;
            if (raw >= eps && raw <= 1.0f - eps)
                g = -(target / (pc + eps) - (1.0f - target) / (1.0f - pc + eps));
            g *= scale * invn;
            dp[i] = accumulate ? dp[i] + g : g;
        }
    }
    __shared__ float red[32];
    acc = warp_sum(acc);
    if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < blockDim.x / 32 ? red[threadIdx.x] : 0.0f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(loss_out, v * invn * scale);
    }
}

// -------------------------------------------------------------------------------------------------
// tf.keras Adam on a flat arena
// -------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ theta, const float* __restrict__ grad,
                            float* __restrict__ m, float* __restrict__ v, int64_t n, float lr_t,
                            float b1, float b2, float eps, float gscale) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float g = grad[i] * gscale;
        const float mi = b1 * m[i] + (1.0f - b1) * g;
        const float vi = b2 * v[i] + (1.0f - b2) * g * g;
        m[i] = mi; v[i] = vi;
        theta[i] -= lr_t * mi / (sqrtf(vi) + eps);
    }
}

// -------------------------------------------------------------------------------------------------
// ConvNextBlock's layer scale (blocks.py:166-179): y = gamma[c] * x with a trainable per-channel gamma.
// bwd: dx = gamma[c] * dy, dgamma[c] += sum over pixels of dy * x.
// -------------------------------------------------------------------------------------------------
__global__ void channel_scale_fwd_kernel(const float* __restrict__ x, int x_ld, const float* __restrict__ gamma,
                                         float* __restrict__ y, int y_ld, int64_t n_pix, int C) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int64_t total = n_pix * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / C;
        const int c = (int)(i - p * C);
        y[p * y_ld + c] = __ldg(gamma + c) * __ldg(x + p * x_ld + c);
    }
}

// block = 256 threads = (256 / C') pixel lanes x C' channel lanes, C' = C rounded up to a power of two <= 256
__global__ void __launch_bounds__(256) channel_scale_bwd_kernel(const float* __restrict__ dy, int dy_ld,
                                                                const float* __restrict__ x, int x_ld,
                                                                const float* __restrict__ gamma, float* __restrict__ dx,
                                                                int dx_ld, float* __restrict__ dgamma, int64_t n_pix,
                                                                int C, int cp) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    __shared__ float red[256];
    const int c = threadIdx.x % cp, lane = threadIdx.x / cp, lanes = 256 / cp;
    float acc = 0.0f;
    if (c < C) {
        const float g = __ldg(gamma + c);
        for (int64_t p = (int64_t)blockIdx.x * lanes + lane; p < n_pix; p += (int64_t)gridDim.x * lanes) {
            const float d = __ldg(dy + p * dy_ld + c);
            acc = fmaf(d, __ldg(x + p * x_ld + c), acc);
            if (dx) dx[p * dx_ld + c] = g * d;
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (lane == 0 && c < C && dgamma) {
        for (int l = 1; l < lanes; ++l) acc += red[l * cp + c];
        atomicAdd(dgamma + c, acc);
    }
}

// -------------------------------------------------------------------------------------------------
// data path: separable resampling with per-output tap tables -- cv2.resize (utils.py:341-401) for every
// interpolation the reference offers (inter_area up/down, nearest, bilinear, bicubic, lanczos4).  The tables hold, per
// output row (column), K source indices and float32 weights taken from cv2 itself (dataloader.DeviceDataGenerator
// resizes an identity matrix), so the arithmetic is cv2's: horizontal pass first, then the vertical one.
// -------------------------------------------------------------------------------------------------
__global__ void resample_taps_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C,
                                     int Ho, int Wo, const int* __restrict__ iy, const float* __restrict__ wy, int Ky,
                                     const int* __restrict__ ix, const float* __restrict__ wx, int Kx, int y_ld,
                                     int y_coff) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int64_t total = (int64_t)N * Ho * Wo * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t t = i / C;
        const int ox = (int)(t % Wo); t /= Wo;
        const int oy = (int)(t % Ho);
        const int n = (int)(t / Ho);
        const float* base = x + (int64_t)n * H * W * C + c;
        float acc = 0.0f;
        for (int j = 0; j < Ky; ++j) {
            const float wj = __ldg(wy + oy * Ky + j);
            if (wj == 0.0f) continue;
            const float* row = base + (int64_t)__ldg(iy + oy * Ky + j) * W * C;
            float h = 0.0f;
            for (int k = 0; k < Kx; ++k) {
                const float wk = __ldg(wx + ox * Kx + k);
                if (wk != 0.0f) h = fmaf(wk, __ldg(row + (int64_t)__ldg(ix + ox * Kx + k) * C), h);
            }
            acc = fmaf(wj, h, acc);
        }
        y[(((int64_t)n * Ho + oy) * Wo + ox) * y_ld + y_coff + c] = acc;
    }
}

// -------------------------------------------------------------------------------------------------
// data path: s x s block mean, summation order of cv2's resizeAreaFast (utils.py:376-384):
// k runs row-major over the s*s window, accumulated as sum += ((v0+v1)+v2)+v3 per group of four,
// remainder one by one, then multiplied by 1/(s*s).
// -------------------------------------------------------------------------------------------------
__global__ void avgpool_coarsen_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H,
                                       int W, int C, int s) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int Ho = H / s, Wo = W / s;
    const int64_t total = (int64_t)N * Ho * Wo * C;
    const int area = s * s;
    const float scale = 1.0f / (float)area;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t t = i / C;
        const int ox = (int)(t % Wo); t /= Wo;
        const int oy = (int)(t % Ho);
        const int n = (int)(t / Ho);
        const float* base = x + (((int64_t)n * H + oy * s) * W + ox * s) * C + c;
        float sum = 0.0f;
        int k = 0;
        for (; k <= area - 4; k += 4) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kk = k + j;
                const int sy = kk / s, sx = kk - sy * s;
                v[j] = __ldg(base + ((int64_t)sy * W + sx) * C);
            }
            sum = __fadd_rn(sum, __fadd_rn(__fadd_rn(__fadd_rn(v[0], v[1]), v[2]), v[3]));
        }
        for (; k < area; ++k) {
            const int sy = k / s, sx = k - sy * s;
            sum = __fadd_rn(sum, __ldg(base + ((int64_t)sy * W + sx) * C));
        }
        y[i] = __fmul_rn(sum, scale);
    }
}

// -------------------------------------------------------------------------------------------------
// bilinear resize, half-pixel centres, edge clamp (tf.image.resize / keras Resizing)
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilinear_coords(int o, int in, int out, int& i0, int& i1, float& f) {
    const float scale = (float)in / (float)out;
    const float s = ((float)o + 0.5f) * scale - 0.5f;
    const float fl = floorf(s);
    f = s - fl;
    const int b = (int)fl;
    i0 = min(max(b, 0), in - 1);
    i1 = min(max(b + 1, 0), in - 1);
}

// -------------------------------------------------------------------------------------------------
// nearest / bicubic resize as Keras Resizing(interpolation=...) = tf.image.resize(method, antialias=False) runs them
// (TF kernels ResizeNearestNeighbor / ResizeBicubic with half_pixel_centers=True; restated from the published source):
//   nearest: in = min(floor((out + 0.5) * scale), in - 1)
//   bicubic: Keys cubic, A = -0.5; in_loc = floor((out + 0.5) * scale - 0.5); the fractional offset is quantised to
//            1/1024 (TF's coefficient table); taps whose index leaves the image get weight 0 and the remaining
//            weights are renormalised to sum 1
// Both are separable: up to 4 taps per axis.
// -------------------------------------------------------------------------------------------------
struct Taps { int idx[4]; float w[4]; int n; };

__device__ __forceinline__ Taps resize_taps(int o, int in, int out, int method) {
    Taps t;
    const float scale = (float)in / (float)out;
    if (method == 1) {
        t.n = 1;
        t.idx[0] = min((int)floorf(((float)o + 0.5f) * scale), in - 1);
        t.w[0] = 1.0f;
        return t;
    }
    const float A = -0.5f;
    const float loc = ((float)o + 0.5f) * scale - 0.5f;
    const float fl = floorf(loc);
    const int b = (int)fl;
    const int off = (int)lrintf((loc - fl) * 1024.0f);
    const float x0 = (float)off / 1024.0f, x1 = (float)(1024 - off) / 1024.0f;
    // coefficient table: [2i] = ((A+2) x - (A+3)) x^2 + 1 at x = i/1024; [2i+1] = ((A x - 5A) x + 8A) x - 4A at x + 1
    const float near0 = ((A + 2.0f) * x0 - (A + 3.0f)) * x0 * x0 + 1.0f;
    const float far0 = ((A * (x0 + 1.0f) - 5.0f * A) * (x0 + 1.0f) + 8.0f * A) * (x0 + 1.0f) - 4.0f * A;
    const float near1 = ((A + 2.0f) * x1 - (A + 3.0f)) * x1 * x1 + 1.0f;
    const float far1 = ((A * (x1 + 1.0f) - 5.0f * A) * (x1 + 1.0f) + 8.0f * A) * (x1 + 1.0f) - 4.0f * A;
    const float wraw[4] = {far0, near0, near1, far1};
    float sum = 0.0f;
    t.n = 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int want = b - 1 + k;
        const int got = min(max(want, 0), in - 1);
        t.idx[k] = got;
        t.w[k] = got == want ? wraw[k] : 0.0f;
        sum += t.w[k];
    }
    if (fabsf(sum) >= 1000.0f * 1.17549435e-38f) {
        const float inv = 1.0f / sum;
#pragma unroll
        for (int k = 0; k < 4; ++k) t.w[k] *= inv;
    }
    return t;
}

// fwd: y[o] = sum w x[taps];  bwd (transpose = 1): scatter g * w onto dx with atomics
__global__ void resize_taps_kernel(const float* __restrict__ src, int src_ld, float* __restrict__ dst, int dst_ld,
                                   int N, int H, int W, int C, int Ho, int Wo, int method, int transpose) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int64_t total = (int64_t)N * Ho * Wo * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t t = i / C;
        const int ox = (int)(t % Wo); t /= Wo;
        const int oy = (int)(t % Ho);
        const int n = (int)(t / Ho);
        const Taps ty = resize_taps(oy, H, Ho, method), tx = resize_taps(ox, W, Wo, method);
        const int64_t o_hi = (((int64_t)n * Ho + oy) * Wo + ox);
        if (!transpose) {
            const float* b = src + (int64_t)n * H * W * src_ld + c;
            float acc = 0.0f;
            for (int a = 0; a < ty.n; ++a) {
                float row = 0.0f;
                for (int k = 0; k < tx.n; ++k) row = fmaf(tx.w[k], __ldg(b + ((int64_t)ty.idx[a] * W + tx.idx[k]) * src_ld), row);
                acc = fmaf(ty.w[a], row, acc);
            }
            dst[o_hi * dst_ld + c] = acc;
        } else {
            const float g = __ldg(src + o_hi * src_ld + c);
            float* b = dst + (int64_t)n * H * W * dst_ld + c;
            for (int a = 0; a < ty.n; ++a)
                for (int k = 0; k < tx.n; ++k) {
                    const float wgt = ty.w[a] * tx.w[k];
                    if (wgt != 0.0f) atomicAdd(b + ((int64_t)ty.idx[a] * W + tx.idx[k]) * dst_ld, g * wgt);
                }
        }
    }
}

__global__ void resize_bilinear_fwd_kernel(const float* __restrict__ x, int x_ld, float* __restrict__ y,
                                           int y_ld, int N, int H, int W, int C, int Ho, int Wo) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int64_t total = (int64_t)N * Ho * Wo * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t t = i / C;
        const int ox = (int)(t % Wo); t /= Wo;
        const int oy = (int)(t % Ho);
        const int n = (int)(t / Ho);
        int y0, y1, x0, x1; float fy, fx;
        bilinear_coords(oy, H, Ho, y0, y1, fy);
        bilinear_coords(ox, W, Wo, x0, x1, fx);
        const float* b = x + (int64_t)n * H * W * x_ld + c;
        const float v00 = __ldg(b + ((int64_t)y0 * W + x0) * x_ld), v01 = __ldg(b + ((int64_t)y0 * W + x1) * x_ld);
        const float v10 = __ldg(b + ((int64_t)y1 * W + x0) * x_ld), v11 = __ldg(b + ((int64_t)y1 * W + x1) * x_ld);
        const float top = v00 + (v01 - v00) * fx, bot = v10 + (v11 - v10) * fx;
        y[(((int64_t)n * Ho + oy) * Wo + ox) * y_ld + c] = top + (bot - top) * fy;
    }
}

__global__ void resize_bilinear_bwd_kernel(const float* __restrict__ dy, int dy_ld, float* __restrict__ dx,
                                           int dx_ld, int N, int H, int W, int C, int Ho, int Wo) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int64_t total = (int64_t)N * Ho * Wo * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t t = i / C;
        const int ox = (int)(t % Wo); t /= Wo;
        const int oy = (int)(t % Ho);
        const int n = (int)(t / Ho);
        int y0, y1, x0, x1; float fy, fx;
        bilinear_coords(oy, H, Ho, y0, y1, fy);
        bilinear_coords(ox, W, Wo, x0, x1, fx);
        const float g = __ldg(dy + (((int64_t)n * Ho + oy) * Wo + ox) * dy_ld + c);
        float* b = dx + (int64_t)n * H * W * dx_ld + c;
        atomicAdd(b + ((int64_t)y0 * W + x0) * dx_ld, g * (1.0f - fy) * (1.0f - fx));
        atomicAdd(b + ((int64_t)y0 * W + x1) * dx_ld, g * (1.0f - fy) * fx);
        atomicAdd(b + ((int64_t)y1 * W + x0) * dx_ld, g * fy * (1.0f - fx));
        atomicAdd(b + ((int64_t)y1 * W + x1) * dx_ld, g * fy * fx);
    }
}

// -------------------------------------------------------------------------------------------------
// 2x2 max-pool (stride 2, floor)
// -------------------------------------------------------------------------------------------------
__global__ void maxpool2_fwd_kernel(const float* __restrict__ x, int x_ld, float* __restrict__ y, int y_ld,
                                    int N, int H, int W, int C) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int Ho = H / 2, Wo = W / 2;
    const int64_t total = (int64_t)N * Ho * Wo * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t t = i / C;
        const int ox = (int)(t % Wo); t /= Wo;
        const int oy = (int)(t % Ho);
        const int n = (int)(t / Ho);
        const float* b = x + (((int64_t)n * H + 2 * oy) * W + 2 * ox) * x_ld + c;
        const float m = fmaxf(fmaxf(__ldg(b), __ldg(b + x_ld)),
                              fmaxf(__ldg(b + (int64_t)W * x_ld), __ldg(b + (int64_t)(W + 1) * x_ld)));
        y[(((int64_t)n * Ho + oy) * Wo + ox) * y_ld + c] = m;
    }
}

// one thread per INPUT element: writes dx fully (first max in row-major window order gets dy)
__global__ void maxpool2_bwd_kernel(const float* __restrict__ x, int x_ld, const float* __restrict__ dy,
                                    int dy_ld, float* __restrict__ dx, int dx_ld, int N, int H, int W,
                                    int C) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int Ho = H / 2, Wo = W / 2;
    const int64_t total = (int64_t)N * H * W * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t t = i / C;
        const int ix = (int)(t % W); t /= W;
        const int iy = (int)(t % H);
        const int n = (int)(t / H);
        const int oy = iy / 2, ox = ix / 2;
        float g = 0.0f;
        if (oy < Ho && ox < Wo) {
            const float* b = x + (((int64_t)n * H + 2 * oy) * W + 2 * ox) * x_ld + c;
            const float v[4] = {__ldg(b), __ldg(b + x_ld), __ldg(b + (int64_t)W * x_ld),
                                __ldg(b + (int64_t)(W + 1) * x_ld)};
            int arg = 0;
            float m = v[0];
#pragma unroll
            for (int k = 1; k < 4; ++k)
                if (v[k] > m) { m = v[k]; arg = k; }
            const int me = (iy - 2 * oy) * 2 + (ix - 2 * ox);
            if (me == arg) g = __ldg(dy + (((int64_t)n * Ho + oy) * Wo + ox) * dy_ld + c);
        }
        dx[(((int64_t)n * H + iy) * W + ix) * dx_ld + c] = g;
    }
}

// -------------------------------------------------------------------------------------------------
// LocallyConnected2D 1x1: y[n,h,w,o] = sum_c x[n,h,w,c] W[h,w,c,o] + b[h,w,o]
// -------------------------------------------------------------------------------------------------
__global__ void local_conv_fwd_kernel(const float* __restrict__ x, int x_ld, const float* __restrict__ w,
                                      const float* __restrict__ b, float* __restrict__ y, int y_ld,
                                      int N, int64_t HW, int Cin, int F) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int64_t total = (int64_t)N * HW * F;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int o = (int)(i % F);
        const int64_t p = i / F;
        const int64_t hw = p % HW;
        float s = b ? __ldg(b + hw * F + o) : 0.0f;
        for (int c = 0; c < Cin; ++c) s += __ldg(x + p * x_ld + c) * __ldg(w + (hw * Cin + c) * F + o);
        y[p * y_ld + o] = s;
    }
}

// one thread per (hw, c): loops over the batch -> deterministic parameter gradients, no atomics
__global__ void local_conv_bwd_kernel(const float* __restrict__ x, int x_ld, const float* __restrict__ dy,
                                      int dy_ld, const float* __restrict__ w, float* __restrict__ dx,
                                      int dx_ld, float* __restrict__ dw, float* __restrict__ db, int N,
                                      int64_t HW, int Cin, int F) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int64_t total = HW * Cin;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cin);
        const int64_t hw = i / Cin;
        float dwacc[8];
        for (int o = 0; o < F && o < 8; ++o) dwacc[o] = 0.0f;
        for (int n = 0; n < N; ++n) {
            const int64_t p = (int64_t)n * HW + hw;
            const float xv = __ldg(x + p * x_ld + c);
            float dxv = 0.0f;
            for (int o = 0; o < F; ++o) {
                const float g = __ldg(dy + p * dy_ld + o);
                dxv += g * __ldg(w + (hw * Cin + c) * F + o);
                if (o < 8) dwacc[o] += xv * g;
            }
            if (dx) dx[p * dx_ld + c] = dxv;
        }
        for (int o = 0; o < F && o < 8; ++o) dw[(hw * Cin + c) * F + o] += dwacc[o];
        if (c == 0 && db) {
            for (int o = 0; o < F; ++o) {
                float s = 0.0f;
                for (int n = 0; n < N; ++n) s += __ldg(dy + ((int64_t)n * HW + hw) * dy_ld + o);
                db[hw * F + o] += s;
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------
// ConvLSTM2D gate math (Keras 2.x: tanh / hard_sigmoid, gate order i,f,c,o)
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ float hard_sigmoid(float x) { return fminf(fmaxf(0.2f * x + 0.5f, 0.0f), 1.0f); }
__device__ __forceinline__ float hard_sigmoid_grad(float x) {
    const float t = 0.2f * x + 0.5f;
    return (t > 0.0f && t < 1.0f) ? 0.2f : 0.0f;
}

__global__ void convlstm_gates_fwd_kernel(const float* __restrict__ z, const float* __restrict__ c_prev,
                                          float* __restrict__ c, float* __restrict__ h, int h_ld,
                                          float* __restrict__ gates, int64_t n_pix, int F) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int64_t total = n_pix * F;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / F;
        const int k = (int)(i - p * F);
        const float* zp = z + p * 4 * F;
        const float zi = zp[k], zf = zp[F + k], zg = zp[2 * F + k], zo = zp[3 * F + k];
        const float gi = hard_sigmoid(zi), gf = hard_sigmoid(zf), gg = tanhf(zg), go = hard_sigmoid(zo);
        const float cp = c_prev ? c_prev[i] : 0.0f;
        const float cn = gf * cp + gi * gg;
        c[i] = cn;
        h[p * h_ld + k] = go * tanhf(cn);
        float* gp = gates + p * 4 * F;
        // store the activated gates; the sign of the hard-sigmoid slope is recoverable from them
        gp[k] = gi; gp[F + k] = gf; gp[2 * F + k] = gg; gp[3 * F + k] = go;
    }
}

__global__ void convlstm_gates_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev,
                                          const float* __restrict__ c, const float* __restrict__ dh,
                                          int dh_ld, const float* __restrict__ dc_next,
                                          float* __restrict__ dz, float* __restrict__ dc_prev,
                                          int64_t n_pix, int F) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int64_t total = n_pix * F;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / F;
        const int k = (int)(i - p * F);
        const float* gp = gates + p * 4 * F;
        const float gi = gp[k], gf = gp[F + k], gg = gp[2 * F + k], go = gp[3 * F + k];
        const float tc = tanhf(c[i]);
        const float dhv = dh[p * dh_ld + k];
        float dc = dhv * go * (1.0f - tc * tc);
        if (dc_next) dc += dc_next[i];
        const float cp = c_prev ? c_prev[i] : 0.0f;
        // hard-sigmoid slope is 0.2 strictly inside (0,1)
        const float si = (gi > 0.0f && gi < 1.0f) ? 0.2f : 0.0f;
        const float sf = (gf > 0.0f && gf < 1.0f) ? 0.2f : 0.0f;
        const float so = (go > 0.0f && go < 1.0f) ? 0.2f : 0.0f;
        float* dzp = dz + p * 4 * F;
        dzp[k] = dc * gg * si;
        dzp[F + k] = dc * cp * sf;
        dzp[2 * F + k] = dc * gi * (1.0f - gg * gg);
        dzp[3 * F + k] = dhv * tc * so;
        dc_prev[i] = dc * gf;
    }
}

__global__ void adam_dev_kernel(float* __restrict__ theta, const float* __restrict__ grad,
                                float* __restrict__ m, float* __restrict__ v, int64_t n,
                                const float* __restrict__ lr_t_dev, float b1, float b2, float eps,
                                float gscale) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const float lr_t = __ldg(lr_t_dev);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float g = grad[i] * gscale;
        const float mi = b1 * m[i] + (1.0f - b1) * g;
        const float vi = b2 * v[i] + (1.0f - b2) * g * g;
        m[i] = mi; v[i] = vi;
        theta[i] -= lr_t * mi / (sqrtf(vi) + eps);
    }
}

// dst[b][a][:] = src[a][b][:]  (frame = contiguous run of `fe` floats, fe % 4 == 0 -> float4 path)
__global__ void permute_frames_kernel(const float* __restrict__ src, float* __restrict__ dst, int A, int B,
                                      int64_t fe) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int64_t total = (int64_t)A * B * fe;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i % fe;
        const int64_t ab = i / fe;
        const int b = (int)(ab % B), a = (int)(ab / B);
        dst[((int64_t)b * A + a) * fe + e] = __ldg(src + i);
    }
}

// zero-pad bottom/right: dst (N,Hd,Wd,C) <- src (N,Hs,Ws,C);  crop=1 is the adjoint (dst <- src window)
__global__ void pad_br_kernel(const float* __restrict__ src, int src_ld, float* __restrict__ dst, int dst_ld,
                              int N, int Hs, int Ws, int Hd, int Wd, int C) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int64_t total = (int64_t)N * Hd * Wd * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t t = i / C;
        const int x = (int)(t % Wd); t /= Wd;
        const int y = (int)(t % Hd);
        const int n = (int)(t / Hd);
        float v = 0.0f;
        if (y < Hs && x < Ws) v = __ldg(src + (((int64_t)n * Hs + y) * Ws + x) * src_ld + c);
        dst[(((int64_t)n * Hd + y) * Wd + x) * dst_ld + c] = v;
    }
}

template <typename K, typename... Args>
static int launch1d(const char* name, K kernel, int64_t work, cudaStream_t st, Args... args) {
    if (work <= 0) return DL4DS_OK;
    kernel<<<grid_for(work, 256), 256, 0, st>>>(args...);
    return check_launch(name);
}

static void pick_xy(int C, int& TX, int& KS) {
    TX = 1;
    while (TX < C && TX < 64) TX <<= 1;
    KS = (int)cdiv(C, TX);
}


// -------------------------------------------------------------------------------------------------
// Composition of a linear convolution + depth_to_space(r) with the 1x1 convolution that follows it
// (the last x2 stage of SubpixelConvolutionBlock, blocks.py:421-427, feeding TransitionLast,
// sp_postups.py:205 / blocks.py:299): a 1x1 conv after depth_to_space applies the same Cm x Co matrix
// to each of the r*r sub-pixel channel groups, so
//   W_eff[row][d*Co + co] = sum_c W1[row][d*Cm + c] * W2[c][co]          row = (tap, ci), d < r*r
//   b_eff[d*Co + co]      = sum_c b1[d*Cm + c] * W2[c][co] + b2[co]
// and the Cm-channel HR tensor (201 MB per step for the headline model) never exists.  The backward
// kernel is the exact chain rule from (dW_eff, db_eff) to the four original parameter gradients.
// -------------------------------------------------------------------------------------------------
__global__ void spc_compose_kernel(const float* __restrict__ w1, const float* __restrict__ b1,
                                   const float* __restrict__ w2, const float* __restrict__ b2,
                                   float* __restrict__ weff, float* __restrict__ beff, int rows, int Cm, int Co, int R2) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int ne = R2 * Co;
    const int total = (rows + 1) * ne;                       // row == rows: the bias row
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int row = i / ne, e = i - row * ne;
        const int d = e / Co, co = e - d * Co;
        const float* src = (row < rows ? w1 + (int64_t)row * R2 * Cm : b1) + d * Cm;
        float acc = 0.0f;
        for (int c = 0; c < Cm; ++c) acc = fmaf(__ldg(src + c), __ldg(w2 + c * Co + co), acc);
        if (row < rows) weff[(int64_t)row * ne + e] = acc;
        else beff[e] = acc + (b2 ? __ldg(b2 + co) : 0.0f);
    }
}

// gW1[row][d*Cm + c] += sum_co dWeff[row][d*Co + co] * W2[c][co]   (row == rows: gb1 from dbeff)
__global__ void spc_chain_w1_kernel(const float* __restrict__ dweff, const float* __restrict__ dbeff,
                                    const float* __restrict__ w2, float* __restrict__ gw1, float* __restrict__ gb1,
                                    int rows, int Cm, int Co, int R2) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int n1 = R2 * Cm, ne = R2 * Co;
    const int total = (rows + 1) * n1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int row = i / n1, e = i - row * n1;
        const int d = e / Cm, c = e - d * Cm;
        const float* src = (row < rows ? dweff + (int64_t)row * ne : dbeff) + d * Co;
        float acc = 0.0f;
        for (int co = 0; co < Co; ++co) acc = fmaf(__ldg(src + co), __ldg(w2 + c * Co + co), acc);
        if (row < rows) gw1[(int64_t)row * n1 + e] += acc;
        else if (gb1) gb1[e] += acc;
    }
}

// gW2[c][co] += sum_{row,d} W1[row][d*Cm+c] * dWeff[row][d*Co+co] + sum_d b1[d*Cm+c] * dbeff[d*Co+co]
// gb2[co]    += sum_d dbeff[d*Co+co]                     one block per (c, co): warp-shuffle + smem reduction
__global__ void __launch_bounds__(256) spc_chain_w2_kernel(const float* __restrict__ w1, const float* __restrict__ b1,
                                                           const float* __restrict__ dweff, const float* __restrict__ dbeff,
                                                           float* __restrict__ gw2, float* __restrict__ gb2,
                                                           int rows, int Cm, int Co, int R2) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    __shared__ float red[8];
    const int c = blockIdx.x / Co, co = blockIdx.x - c * Co;
    const int n1 = R2 * Cm, ne = R2 * Co;
    float acc = 0.0f;
    for (int i = threadIdx.x; i < (rows + 1) * R2; i += 256) {
        const int row = i / R2, d = i - row * R2;
        const float a = row < rows ? __ldg(w1 + (int64_t)row * n1 + d * Cm + c) : (b1 ? __ldg(b1 + d * Cm + c) : 0.0f);
        const float g = row < rows ? __ldg(dweff + (int64_t)row * ne + d * Co + co) : __ldg(dbeff + d * Co + co);
        acc = fmaf(a, g, acc);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int w = 0; w < 8; ++w) t += red[w];
        gw2[c * Co + co] += t;
        if (c == 0 && gb2) {
            float sb = 0.0f;
            for (int d = 0; d < R2; ++d) sb += __ldg(dbeff + d * Co + co);
            gb2[co] += sb;
        }
    }
}


// -------------------------------------------------------------------------------------------------
// device-resident batch assembly (dataloader.py:297-360 without the per-sample host loop): sample
// picks[b] of the resident array, cropped at (y0[b], x0[b]) to (ph, pw), written into channels
// [dst_coff, dst_coff + C) of the batch tensor.  idx / y0 / x0 are DEVICE int32 arrays (y0, x0 may be NULL).
// -------------------------------------------------------------------------------------------------
__global__ void gather_crop_kernel(const float* __restrict__ src, const int* __restrict__ idx,
                                   const int* __restrict__ y0, const int* __restrict__ x0, float* __restrict__ dst,
                                   int n, int H, int W, int C, int ph, int pw, int dst_ld, int dst_coff) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int64_t total = (int64_t)n * ph * pw * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t t = i / C;
        const int x = (int)(t % pw); t /= pw;
        const int y = (int)(t % ph);
        const int b = (int)(t / ph);
        const int sy = y + (y0 ? __ldg(y0 + b) : 0), sx = x + (x0 ? __ldg(x0 + b) : 0);
        dst[(((int64_t)b * ph + y) * pw + x) * dst_ld + dst_coff + c] =
            __ldg(src + (((int64_t)__ldg(idx + b) * H + sy) * W + sx) * C + c);
    }
}

// -------------------------------------------------------------------------------------------------
// Conv2DTranspose(k, strides=s, 'same') (DeconvolutionBlock, blocks.py:508-516) as a stride-1 convolution on the
// input grid followed by depth_to_space(s): output pixel (s*a + dy, s*b + dx) only sees the kernel taps kh with
// (dy + kh - pad) % s == 0, at input row a + (dy + kh - pad) / s.  Collecting the s*s phases as output-channel
// groups gives a Kp x Kp convolution (Kp = ceil-ish(k/s): 5 for k=9, s=2) to s*s*Co channels in HWIO layout:
//   wp[my][mx][ci][(dy*s + dx)*Co + co] = w[k-1-kh][k-1-kw][co][ci]   (Keras layout (kh,kw,Co,Ci), flipped), or 0
// with kh = s*(my + off_min) - dy + pad, kw likewise -- which runs on the tensor-core kernels with the fused
// depth_to_space store instead of the CUDA-core fractional-stride path.  backward != 0 scatters dwp back: dw += .
// -------------------------------------------------------------------------------------------------
__global__ void convt_rearrange_kernel(float* __restrict__ w, float* __restrict__ wp, int k, int s, int pad, int off_min,
                                       int Kp, int Co, int Ci, int backward) {
    pdl_launch_dependents();    // lets a PDL-launched successor start its prologue (common.cuh); this kernel itself is launched normally
    const int ne = s * s * Co;
    const int64_t total = (int64_t)Kp * Kp * Ci * ne;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(i % ne);
        int64_t r = i / ne;
        const int ci = (int)(r % Ci); r /= Ci;
        const int mx = (int)(r % Kp);
        const int my = (int)(r / Kp);
        const int d = e / Co, co = e - d * Co;
        const int dy = d / s, dx = d - dy * s;
        const int kh = s * (my + off_min) - dy + pad, kw = s * (mx + off_min) - dx + pad;
        const bool ok = kh >= 0 && kh < k && kw >= 0 && kw < k;
        const int64_t wi = ok ? ((int64_t)((k - 1 - kh) * k + (k - 1 - kw)) * Co + co) * Ci + ci : 0;
        if (backward) {
            if (ok) w[wi] += wp[i];          // every (kh, kw, co, ci) has exactly one image: no atomics needed
        } else {
            wp[i] = ok ? w[wi] : 0.0f;
        }
    }
}
}  // namespace dl4ds

using namespace dl4ds;

extern "C" {

int dl4ds_spc_pointwise_compose(const float* w1, const float* b1, const float* w2, const float* b2,
                                float* weff, float* beff, int rows, int Cm, int Co, int r, void* stream) {
    DL4DS_REQUIRE(w1 && b1 && w2 && weff && beff, DL4DS_E_BADARG, "spc_pointwise_compose: null pointer");
    DL4DS_REQUIRE(rows > 0 && Cm > 0 && Co > 0 && r > 1, DL4DS_E_SHAPE, "spc_pointwise_compose: bad shape");
    const int total = (rows + 1) * r * r * Co;
    spc_compose_kernel<<<(unsigned)cdiv(total, 256), 256, 0, as_stream(stream)>>>(w1, b1, w2, b2, weff, beff, rows, Cm, Co, r * r);
    return check_launch("spc_compose");
}

int dl4ds_spc_pointwise_chain(const float* w1, const float* b1, const float* w2, const float* dweff, const float* dbeff,
                              float* gw1, float* gb1, float* gw2, float* gb2, int rows, int Cm, int Co, int r,
                              void* stream) {
    DL4DS_REQUIRE(w1 && w2 && dweff && dbeff && gw1 && gw2, DL4DS_E_BADARG, "spc_pointwise_chain: null pointer");
    DL4DS_REQUIRE(rows > 0 && Cm > 0 && Co > 0 && r > 1, DL4DS_E_SHAPE, "spc_pointwise_chain: bad shape");
    cudaStream_t st = as_stream(stream);
    const int total = (rows + 1) * r * r * Cm;
    spc_chain_w1_kernel<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(dweff, dbeff, w2, gw1, gb1, rows, Cm, Co, r * r);
    int rc = check_launch("spc_chain_w1");
    if (rc) return rc;
    spc_chain_w2_kernel<<<(unsigned)(Cm * Co), 256, 0, st>>>(w1, b1, dweff, dbeff, gw2, gb2, rows, Cm, Co, r * r);
    return check_launch("spc_chain_w2");
}

int dl4ds_bias_act_bwd(const float* dy, int dy_ld, const float* y, int y_ld, float* dz, int dz_ld,
                       float* dbias, int N, int Ho, int Wo, int C, int act, int d2s_r, void* stream) {
    DL4DS_REQUIRE(dy, DL4DS_E_BADARG, "bias_act_bwd: null dy");
    DL4DS_REQUIRE(N > 0 && Ho > 0 && Wo > 0 && C > 0, DL4DS_E_SHAPE, "bias_act_bwd: bad shape");
    DL4DS_REQUIRE(act >= 0 && act <= 3, DL4DS_E_BADARG, "bias_act_bwd: act");
    if (d2s_r <= 1) d2s_r = 1;
    if (d2s_r > 1) {
        DL4DS_REQUIRE(C % (d2s_r * d2s_r) == 0, DL4DS_E_BADARG, "bias_act_bwd: d2s needs C %% r^2 == 0");
        DL4DS_REQUIRE(act == DL4DS_ACT_NONE || y != nullptr, DL4DS_E_BADARG, "bias_act_bwd: y needed");
        DL4DS_REQUIRE(dz != nullptr && dz != dy, DL4DS_E_BADARG, "bias_act_bwd: d2s needs distinct dz");
    } else {
        DL4DS_REQUIRE(act == DL4DS_ACT_NONE || y != nullptr, DL4DS_E_BADARG, "bias_act_bwd: y needed");
    }
    int TX, KS;
    const int64_t n_pix = (int64_t)N * Ho * Wo;
    cudaStream_t st = as_stream(stream);
    {
        int rc = bias_act_bwd_vec4(dy, dy_ld, y, y_ld, dz, dz_ld, dbias, n_pix, Ho, Wo, C, act, d2s_r, st);
        if (rc != DL4DS_E_UNSUPPORTED) return rc;
    }
    pick_xy(C, TX, KS);
    DL4DS_REQUIRE(KS <= 8, DL4DS_E_UNSUPPORTED, "bias_act_bwd: C > 512 unsupported");
    const int PY = 256 / TX;
    dim3 block(TX, PY);
    const int grid = grid_for(n_pix, PY * 8, 4 * kNumSMs);
#define LAUNCH_BAB(K) bias_act_bwd_kernel<K><<<grid, block, 0, st>>>(dy, dy_ld, y, y_ld, dz, dz_ld, dbias, n_pix, Ho, Wo, C, act, d2s_r)
    if (KS == 1) LAUNCH_BAB(1);
    else if (KS == 2) LAUNCH_BAB(2);
    else if (KS <= 4) LAUNCH_BAB(4);
    else LAUNCH_BAB(8);
#undef LAUNCH_BAB
    return check_launch("bias_act_bwd");
}

int dl4ds_add(const float* a, int a_ld, const float* b, int b_ld, float* out, int out_ld,
              int64_t n_pix, int C, int act, void* stream) {
    DL4DS_REQUIRE(a && b && out, DL4DS_E_BADARG, "add: null pointer");
    DL4DS_REQUIRE(n_pix >= 0 && C > 0, DL4DS_E_SHAPE, "add: bad shape");
    cudaStream_t st = as_stream(stream);
    if (a_ld == C && b_ld == C && out_ld == C && (n_pix * C) % 4 == 0 && aligned16(a) && aligned16(b) &&
        aligned16(out)) {
        const int64_t n4 = n_pix * C / 4;
        return launch1d("add", add_vec4_kernel, n4, st, reinterpret_cast<const float4*>(a),
                        reinterpret_cast<const float4*>(b), reinterpret_cast<float4*>(out), n4, act);
    }
    return launch1d("add", add_kernel, n_pix * C, st, a, a_ld, b, b_ld, out, out_ld, n_pix, C, act);
}

int dl4ds_copy_channels(const float* src, int src_ld, float* dst, int dst_ld, int64_t n_pix, int C,
                        int accumulate, void* stream) {
    DL4DS_REQUIRE(src && dst, DL4DS_E_BADARG, "copy_channels: null pointer");
    DL4DS_REQUIRE(n_pix >= 0 && C > 0 && src_ld >= C && dst_ld >= C, DL4DS_E_SHAPE, "copy_channels: bad shape");
    cudaStream_t st = as_stream(stream);
    if (C % 4 == 0 && src_ld % 4 == 0 && dst_ld % 4 == 0 && aligned16(src) && aligned16(dst))
        return launch1d("copy_channels", copy_channels_vec4_kernel, n_pix * (C / 4), st, src, src_ld, dst,
                        dst_ld, n_pix, C / 4, accumulate);
    return launch1d("copy_channels", copy_channels_kernel, n_pix * C, st, src, src_ld, dst, dst_ld, n_pix,
                    C, accumulate);
}

int dl4ds_act_fwd(const float* x, int x_ld, float* y, int y_ld, int64_t n_pix, int C, int act,
                  void* stream) {
    DL4DS_REQUIRE(x && y, DL4DS_E_BADARG, "act_fwd: null pointer");
    return launch1d("act_fwd", act_fwd_kernel, n_pix * C, as_stream(stream), x, x_ld, y, y_ld, n_pix, C, act);
}

int dl4ds_mul(const float* a, const float* b, float* out, int64_t n, void* stream) {
    DL4DS_REQUIRE(a && b && out, DL4DS_E_BADARG, "mul: null pointer");
    return launch1d("mul", mul_kernel, n, as_stream(stream), a, b, out, n);
}

int dl4ds_axpby(float a, const float* x, float b, float* y, int64_t n, void* stream) {
    DL4DS_REQUIRE(x && y, DL4DS_E_BADARG, "axpby: null pointer");
    return launch1d("axpby", axpby_kernel, n, as_stream(stream), a, x, b, y, n);
}

static int group_scale(const char* what, const float* x, int x_ld, float* y, int y_ld, const float* s, const float* dm,
                       float inv, int64_t n_pix, int64_t ppg, int inner, int C, int mode, cudaStream_t st) {
    const int64_t n4 = n_pix * (C / 4);
    if (C % 4 == 0 && y_ld % 4 == 0 && aligned16(y) && (mode == 2 || (x_ld % 4 == 0 && aligned16(x) && aligned16(s))) &&
        (mode == 0 || aligned16(dm)) && n4 < (1ll << 31) && ppg * inner < (1ll << 31)) {
        int64_t blocks = (n4 + 255) / 256;
        if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
        group_scale_vec4_kernel<<<(int)blocks, 256, 0, st>>>(x, x_ld, y, y_ld, s, dm, inv, (int)n4, (int)ppg, inner, C, mode);
        return check_launch(what);
    }
    return launch1d(what, group_scale_kernel, n_pix * C, st, x, x_ld, y, y_ld, s, dm, inv, n_pix, ppg, inner, C, mode);
}

static int group_sum(const float* a, int a_ld, const float* b, int b_ld, float* out, int n_groups,
                     int64_t ppg, int inner, int C, cudaStream_t st) {
    int TX, KS;
    pick_xy(C, TX, KS);
    DL4DS_REQUIRE(KS <= 8, DL4DS_E_UNSUPPORTED, "group_sum: C > 512 unsupported");
    const int PY = 256 / TX;
    int chunks = (int)cdiv(ppg, (int64_t)PY * 16);
    const int max_chunks = (int)cdiv(4 * kNumSMs, n_groups);
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    dim3 grid(chunks, n_groups), block(TX, PY);
    if (cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n_groups * C, st) != cudaSuccess)
        return check_launch("group_sum memset");
    {
        const int G4 = C / 4;
        const bool pow2 = G4 > 0 && (G4 & (G4 - 1)) == 0;
        if (inner == 1 && C % 4 == 0 && pow2 && G4 <= 16 && a_ld % 4 == 0 && aligned16(a) && ppg < (1ll << 30) &&
            (!b || (b_ld % 4 == 0 && aligned16(b)))) {
            const int py = 256 / G4;
            int ch = (int)cdiv(ppg, (int64_t)py * 8);
            const int mc = (int)cdiv(8 * kNumSMs, n_groups);
            if (ch > mc) ch = mc;
            if (ch < 1) ch = 1;
            dim3 g2(ch, n_groups), b2(G4, py);
            if (b) group_sum_vec4_kernel<true><<<g2, b2, 0, st>>>(a, a_ld, b, b_ld, out, (int)ppg, C);
            else group_sum_vec4_kernel<false><<<g2, b2, 0, st>>>(a, a_ld, b, b_ld, out, (int)ppg, C);
            return check_launch("group_sum_vec4");
        }
    }
#define LAUNCH_GS(K)                                                                              \
    do {                                                                                          \
        if (b) group_sum_kernel<K, true><<<grid, block, 0, st>>>(a, a_ld, b, b_ld, out, ppg, inner, C); \
        else group_sum_kernel<K, false><<<grid, block, 0, st>>>(a, a_ld, b, b_ld, out, ppg, inner, C);  \
    } while (0)
    if (KS == 1) LAUNCH_GS(1);
    else if (KS == 2) LAUNCH_GS(2);
    else if (KS <= 4) LAUNCH_GS(4);
    else LAUNCH_GS(8);
#undef LAUNCH_GS
    return check_launch("group_sum");
}

int dl4ds_channel_attention_fwd(const float* x, int x_ld, float* y, int y_ld,
                                const float* w1, const float* b1, const float* w2, const float* b2,
                                float* pooled, float* hidden, float* scale,
                                int n_groups, int64_t pix_per_group, int inner, int C, int Cr,
                                void* stream) {
    DL4DS_REQUIRE(x && y && w1 && b1 && w2 && b2 && pooled && hidden && scale, DL4DS_E_BADARG,
                  "channel_attention_fwd: null pointer");
    DL4DS_REQUIRE(n_groups > 0 && pix_per_group > 0 && inner > 0 && n_groups % inner == 0 && C > 0 && Cr > 0,
                  DL4DS_E_SHAPE, "channel_attention_fwd: bad shape");
    cudaStream_t st = as_stream(stream);
    int rc = group_sum(x, x_ld, nullptr, 0, pooled, n_groups, pix_per_group, inner, C, st);
    if (rc) return rc;
    const float inv = 1.0f / (float)pix_per_group;
    attention_mlp_fwd_kernel<<<(unsigned)cdiv(n_groups, 4), 128, 0, st>>>(pooled, inv, w1, b1, w2, b2, hidden,
                                                                          scale, n_groups, C, Cr);
    rc = check_launch("attention_mlp_fwd");
    if (rc) return rc;
    const int64_t n_pix = (int64_t)n_groups * pix_per_group;
    return group_scale("attention_scale", x, x_ld, y, y_ld, scale, nullptr, 0.0f, n_pix, pix_per_group, inner, C, 0, st);
}

int dl4ds_channel_attention_bwd(const float* x, int x_ld, const float* dy, int dy_ld,
                                float* dx, int dx_ld,
                                const float* w1, const float* w2,
                                const float* pooled, const float* hidden, const float* scale,
                                float* dsum, float* dw1, float* db1, float* dw2, float* db2,
                                int n_groups, int64_t pix_per_group, int inner, int C, int Cr,
                                void* stream) {
    DL4DS_REQUIRE(x && dy && dx && w1 && w2 && pooled && hidden && scale && dsum && dw1 && db1 && dw2 && db2,
                  DL4DS_E_BADARG, "channel_attention_bwd: null pointer");
    DL4DS_REQUIRE(n_groups > 0 && pix_per_group > 0 && inner > 0 && n_groups % inner == 0 && C > 0 && Cr > 0,
                  DL4DS_E_SHAPE, "channel_attention_bwd: bad shape");
    cudaStream_t st = as_stream(stream);
    int rc = group_sum(dy, dy_ld, x, x_ld, dsum, n_groups, pix_per_group, inner, C, st);
    if (rc) return rc;
    const float inv = 1.0f / (float)pix_per_group;
    const size_t smem = 4 * (size_t)(C + Cr) * sizeof(float);
    DL4DS_REQUIRE(smem <= 48 * 1024, DL4DS_E_UNSUPPORTED, "channel_attention_bwd: C too large");
    attention_mlp_bwd_kernel<<<(unsigned)cdiv(n_groups, 4), 128, smem, st>>>(
        pooled, inv, w1, w2, hidden, scale, dsum, dw1, db1, dw2, db2, n_groups, C, Cr);
    rc = check_launch("attention_mlp_bwd");
    if (rc) return rc;
    const int64_t n_pix = (int64_t)n_groups * pix_per_group;
    return group_scale("attention_bwd_scale", dy, dy_ld, dx, dx_ld, scale, dsum, inv, n_pix, pix_per_group, inner, C, 1, st);
}

int dl4ds_group_mean_fwd(const float* x, int x_ld, float* out, int n_groups, int64_t pix_per_group,
                         int C, void* stream) {
    DL4DS_REQUIRE(x && out, DL4DS_E_BADARG, "group_mean_fwd: null pointer");
    DL4DS_REQUIRE(n_groups > 0 && pix_per_group > 0 && C > 0, DL4DS_E_SHAPE, "group_mean_fwd: bad shape");
    cudaStream_t st = as_stream(stream);
    int rc = group_sum(x, x_ld, nullptr, 0, out, n_groups, pix_per_group, 1, C, st);
    if (rc) return rc;
    return launch1d("group_mean_scale", axpby_kernel, (int64_t)n_groups * C, st,
                    1.0f / (float)pix_per_group, (const float*)out, 0.0f, out, (int64_t)n_groups * C);
}

int dl4ds_group_mean_bwd(const float* dout, float* dx, int dx_ld, int n_groups, int64_t pix_per_group,
                         int C, void* stream) {
    DL4DS_REQUIRE(dout && dx, DL4DS_E_BADARG, "group_mean_bwd: null pointer");
    const int64_t n_pix = (int64_t)n_groups * pix_per_group;
    return group_scale("group_mean_bwd", nullptr, 0, dx, dx_ld, nullptr, dout, 1.0f / (float)pix_per_group, n_pix,
                       pix_per_group, 1, C, 2, as_stream(stream));
}

int dl4ds_pixel_loss(const float* y_pred, const float* y_true, float* loss_out, float* dy,
                     int64_t n, int kind, float scale, void* stream) {
    DL4DS_REQUIRE(y_pred && y_true && loss_out, DL4DS_E_BADARG, "pixel_loss: null pointer");
    const int accumulate = (kind & DL4DS_LOSS_ACCUMULATE) ? 1 : 0;
    kind &= ~DL4DS_LOSS_ACCUMULATE;
    DL4DS_REQUIRE(n > 0 && (kind == 0 || kind == 1), DL4DS_E_BADARG, "pixel_loss: bad n/kind");
    pixel_loss_kernel<<<grid_for(n, 256 * 4, 2 * kNumSMs), 256, 0, as_stream(stream)>>>(
        y_pred, y_true, loss_out, dy, n, kind, scale, accumulate);
    return check_launch("pixel_loss");
}

int dl4ds_bce_loss(const float* p, float target, float* loss_out, float* dp, int64_t n,
                   float scale, int accumulate, void* stream) {
    DL4DS_REQUIRE(p && loss_out, DL4DS_E_BADARG, "bce_loss: null pointer");
    DL4DS_REQUIRE(n > 0, DL4DS_E_SHAPE, "bce_loss: n <= 0");
    bce_loss_kernel<<<1, 256, 0, as_stream(stream)>>>(p, target, loss_out, dp, n, scale, accumulate);
    return check_launch("bce_loss");
}

int dl4ds_adam_step(float* theta, const float* grad, float* m, float* v, int64_t n,
                    float lr, float beta1, float beta2, float eps, int t, float grad_scale,
                    void* stream) {
    DL4DS_REQUIRE(theta && grad && m && v, DL4DS_E_BADARG, "adam_step: null pointer");
    DL4DS_REQUIRE(n > 0 && t >= 1, DL4DS_E_BADARG, "adam_step: n <= 0 or t < 1");
    const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)t)) /
                        (1.0 - pow((double)beta1, (double)t));
    return launch1d("adam_step", adam_kernel, n, as_stream(stream), theta, grad, m, v, n, (float)lr_t,
                    beta1, beta2, eps, grad_scale);
}

int dl4ds_adam_step_dev(float* theta, const float* grad, float* m, float* v, int64_t n,
                        const float* lr_t_dev, float beta1, float beta2, float eps, float grad_scale,
                        void* stream) {
    DL4DS_REQUIRE(theta && grad && m && v && lr_t_dev, DL4DS_E_BADARG, "adam_step_dev: null pointer");
    DL4DS_REQUIRE(n > 0, DL4DS_E_BADARG, "adam_step_dev: n <= 0");
    return launch1d("adam_step_dev", adam_dev_kernel, n, as_stream(stream), theta, grad, m, v, n, lr_t_dev,
                    beta1, beta2, eps, grad_scale);
}

int dl4ds_permute_frames(const float* src, float* dst, int A, int B, int64_t frame_elems, void* stream) {
    DL4DS_REQUIRE(src && dst && src != dst, DL4DS_E_BADARG, "permute_frames: null or aliased pointer");
    DL4DS_REQUIRE(A > 0 && B > 0 && frame_elems > 0, DL4DS_E_SHAPE, "permute_frames: bad shape");
    return launch1d("permute_frames", permute_frames_kernel, (int64_t)A * B * frame_elems, as_stream(stream),
                    src, dst, A, B, frame_elems);
}

int dl4ds_pad_bottom_right(const float* src, int src_ld, float* dst, int dst_ld,
                           int N, int Hs, int Ws, int Hd, int Wd, int C, void* stream) {
    DL4DS_REQUIRE(src && dst, DL4DS_E_BADARG, "pad_bottom_right: null pointer");
    DL4DS_REQUIRE(N > 0 && Hs > 0 && Ws > 0 && Hd > 0 && Wd > 0 && C > 0, DL4DS_E_SHAPE,
                  "pad_bottom_right: bad shape");
    return launch1d("pad_bottom_right", pad_br_kernel, (int64_t)N * Hd * Wd * C, as_stream(stream), src,
                    src_ld, dst, dst_ld, N, Hs, Ws, Hd, Wd, C);
}

int dl4ds_gather_crop(const float* src, const int* idx, const int* y0, const int* x0, float* dst,
                      int n, int H, int W, int C, int ph, int pw, int dst_ld, int dst_coff, void* stream) {
    DL4DS_REQUIRE(src && idx && dst, DL4DS_E_BADARG, "gather_crop: null pointer");
    DL4DS_REQUIRE(n > 0 && H > 0 && W > 0 && C > 0 && ph > 0 && pw > 0 && ph <= H && pw <= W, DL4DS_E_SHAPE,
                  "gather_crop: bad shape");
    DL4DS_REQUIRE(dst_ld >= dst_coff + C && dst_coff >= 0, DL4DS_E_SHAPE, "gather_crop: channel slice outside dst_ld");
    return launch1d("gather_crop", gather_crop_kernel, (int64_t)n * ph * pw * C, as_stream(stream), src, idx, y0, x0, dst,
                    n, H, W, C, ph, pw, dst_ld, dst_coff);
}

int dl4ds_convt_rearrange(float* w, float* wp, int k, int stride, int pad, int off_min, int Kp, int Co, int Ci,
                          int backward, void* stream) {
    DL4DS_REQUIRE(w && wp, DL4DS_E_BADARG, "convt_rearrange: null pointer");
    DL4DS_REQUIRE(k > 0 && stride > 1 && Kp > 0 && Co > 0 && Ci > 0, DL4DS_E_SHAPE, "convt_rearrange: bad shape");
    const int64_t total = (int64_t)Kp * Kp * Ci * stride * stride * Co;
    return launch1d("convt_rearrange", convt_rearrange_kernel, total, as_stream(stream), w, wp, k, stride, pad, off_min, Kp,
                    Co, Ci, backward);
}

int dl4ds_avgpool_coarsen(const float* x, float* y, int N, int H, int W, int C, int s, void* stream) {
    DL4DS_REQUIRE(x && y, DL4DS_E_BADARG, "avgpool_coarsen: null pointer");
    DL4DS_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && s > 0 && H % s == 0 && W % s == 0, DL4DS_E_SHAPE,
                  "avgpool_coarsen: H, W must be multiples of s");
    return launch1d("avgpool_coarsen", avgpool_coarsen_kernel, (int64_t)N * (H / s) * (W / s) * C,
                    as_stream(stream), x, y, N, H, W, C, s);
}

int dl4ds_channel_scale_fwd(const float* x, int x_ld, const float* gamma, float* y, int y_ld, int64_t n_pix, int C,
                            void* stream) {
    DL4DS_REQUIRE(x && gamma && y, DL4DS_E_BADARG, "channel_scale_fwd: null pointer");
    DL4DS_REQUIRE(n_pix > 0 && C > 0 && x_ld >= C && y_ld >= C, DL4DS_E_SHAPE, "channel_scale_fwd: bad shape");
    return launch1d("channel_scale_fwd", channel_scale_fwd_kernel, n_pix * C, as_stream(stream), x, x_ld, gamma, y, y_ld,
                    n_pix, C);
}

int dl4ds_channel_scale_bwd(const float* dy, int dy_ld, const float* x, int x_ld, const float* gamma, float* dx,
                            int dx_ld, float* dgamma, int64_t n_pix, int C, void* stream) {
    DL4DS_REQUIRE(dy && x && gamma && (dx || dgamma), DL4DS_E_BADARG, "channel_scale_bwd: null pointer");
    DL4DS_REQUIRE(n_pix > 0 && C > 0 && C <= 256 && dy_ld >= C && x_ld >= C && (!dx || dx_ld >= C), DL4DS_E_SHAPE,
                  "channel_scale_bwd: bad shape (C <= 256)");
    int cp = 1;
    while (cp < C) cp *= 2;
    const int lanes = 256 / cp;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_pix + lanes - 1) / lanes, 8 * kNumSMs));
    channel_scale_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(dy, dy_ld, x, x_ld, gamma, dx, dx_ld, dgamma, n_pix, C, cp);
    return check_launch("channel_scale_bwd");
}

int dl4ds_resample_taps(const float* x, float* y, int N, int H, int W, int C, int Ho, int Wo, const int* iy,
                        const float* wy, int Ky, const int* ix, const float* wx, int Kx, int y_ld, int y_coff,
                        void* stream) {
    DL4DS_REQUIRE(x && y && iy && wy && ix && wx, DL4DS_E_BADARG, "resample_taps: null pointer");
    DL4DS_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0 && Ky > 0 && Kx > 0, DL4DS_E_SHAPE,
                  "resample_taps: bad shape");
    DL4DS_REQUIRE(y_coff >= 0 && y_ld >= y_coff + C, DL4DS_E_SHAPE, "resample_taps: channel slice outside y_ld");
    return launch1d("resample_taps", resample_taps_kernel, (int64_t)N * Ho * Wo * C, as_stream(stream), x, y, N, H, W, C,
                    Ho, Wo, iy, wy, Ky, ix, wx, Kx, y_ld, y_coff);
}

int dl4ds_resize_bilinear_fwd(const float* x, int x_ld, float* y, int y_ld,
                              int N, int H, int W, int C, int Ho, int Wo, void* stream) {
    DL4DS_REQUIRE(x && y, DL4DS_E_BADARG, "resize_bilinear_fwd: null pointer");
    DL4DS_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0, DL4DS_E_SHAPE, "resize_bilinear_fwd: bad shape");
    return launch1d("resize_bilinear_fwd", resize_bilinear_fwd_kernel, (int64_t)N * Ho * Wo * C,
                    as_stream(stream), x, x_ld, y, y_ld, N, H, W, C, Ho, Wo);
}

int dl4ds_resize_bilinear_bwd(const float* dy, int dy_ld, float* dx, int dx_ld,
                              int N, int H, int W, int C, int Ho, int Wo, void* stream) {
    DL4DS_REQUIRE(dy && dx, DL4DS_E_BADARG, "resize_bilinear_bwd: null pointer");
    DL4DS_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0, DL4DS_E_SHAPE, "resize_bilinear_bwd: bad shape");
    return launch1d("resize_bilinear_bwd", resize_bilinear_bwd_kernel, (int64_t)N * Ho * Wo * C,
                    as_stream(stream), dy, dy_ld, dx, dx_ld, N, H, W, C, Ho, Wo);
}

int dl4ds_resize_fwd(const float* x, int x_ld, float* y, int y_ld, int N, int H, int W, int C, int Ho, int Wo,
                     int method, void* stream) {
    DL4DS_REQUIRE(x && y, DL4DS_E_BADARG, "resize_fwd: null pointer");
    DL4DS_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0, DL4DS_E_SHAPE, "resize_fwd: bad shape");
    if (method == DL4DS_RESIZE_BILINEAR) return dl4ds_resize_bilinear_fwd(x, x_ld, y, y_ld, N, H, W, C, Ho, Wo, stream);
    DL4DS_REQUIRE(method == DL4DS_RESIZE_NEAREST || method == DL4DS_RESIZE_BICUBIC, DL4DS_E_UNSUPPORTED,
                  "resize_fwd: method %d not built (bilinear 0, nearest 1, bicubic 2)", method);
    return launch1d("resize_fwd", resize_taps_kernel, (int64_t)N * Ho * Wo * C, as_stream(stream), x, x_ld, y, y_ld, N,
                    H, W, C, Ho, Wo, method, 0);
}

int dl4ds_resize_bwd(const float* dy, int dy_ld, float* dx, int dx_ld, int N, int H, int W, int C, int Ho, int Wo,
                     int method, void* stream) {
    DL4DS_REQUIRE(dy && dx, DL4DS_E_BADARG, "resize_bwd: null pointer");
    DL4DS_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0, DL4DS_E_SHAPE, "resize_bwd: bad shape");
    if (method == DL4DS_RESIZE_BILINEAR) return dl4ds_resize_bilinear_bwd(dy, dy_ld, dx, dx_ld, N, H, W, C, Ho, Wo, stream);
    DL4DS_REQUIRE(method == DL4DS_RESIZE_NEAREST || method == DL4DS_RESIZE_BICUBIC, DL4DS_E_UNSUPPORTED,
                  "resize_bwd: method %d not built (bilinear 0, nearest 1, bicubic 2)", method);
    return launch1d("resize_bwd", resize_taps_kernel, (int64_t)N * Ho * Wo * C, as_stream(stream), dy, dy_ld, dx, dx_ld,
                    N, H, W, C, Ho, Wo, method, 1);
}


int dl4ds_maxpool2_fwd(const float* x, int x_ld, float* y, int y_ld, int N, int H, int W, int C,
                       void* stream) {
    DL4DS_REQUIRE(x && y, DL4DS_E_BADARG, "maxpool2_fwd: null pointer");
    DL4DS_REQUIRE(N > 0 && H >= 2 && W >= 2 && C > 0, DL4DS_E_SHAPE, "maxpool2_fwd: bad shape");
    return launch1d("maxpool2_fwd", maxpool2_fwd_kernel, (int64_t)N * (H / 2) * (W / 2) * C,
                    as_stream(stream), x, x_ld, y, y_ld, N, H, W, C);
}

int dl4ds_maxpool2_bwd(const float* x, int x_ld, const float* dy, int dy_ld, float* dx, int dx_ld,
                       int N, int H, int W, int C, void* stream) {
    DL4DS_REQUIRE(x && dy && dx, DL4DS_E_BADARG, "maxpool2_bwd: null pointer");
    DL4DS_REQUIRE(N > 0 && H >= 2 && W >= 2 && C > 0, DL4DS_E_SHAPE, "maxpool2_bwd: bad shape");
    return launch1d("maxpool2_bwd", maxpool2_bwd_kernel, (int64_t)N * H * W * C, as_stream(stream), x, x_ld,
                    dy, dy_ld, dx, dx_ld, N, H, W, C);
}

int dl4ds_local_conv1x1_fwd(const float* x, int x_ld, const float* w, const float* b,
                            float* y, int y_ld, int N, int H, int W, int Cin, int F, void* stream) {
    DL4DS_REQUIRE(x && w && y, DL4DS_E_BADARG, "local_conv1x1_fwd: null pointer");
    DL4DS_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && F > 0, DL4DS_E_SHAPE, "local_conv1x1_fwd: bad shape");
    return launch1d("local_conv1x1_fwd", local_conv_fwd_kernel, (int64_t)N * H * W * F, as_stream(stream), x,
                    x_ld, w, b, y, y_ld, N, (int64_t)H * W, Cin, F);
}

int dl4ds_local_conv1x1_bwd(const float* x, int x_ld, const float* dy, int dy_ld, const float* w,
                            float* dx, int dx_ld, float* dw, float* db,
                            int N, int H, int W, int Cin, int F, void* stream) {
    DL4DS_REQUIRE(x && dy && w && dw, DL4DS_E_BADARG, "local_conv1x1_bwd: null pointer");
    DL4DS_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && F > 0 && F <= 8, DL4DS_E_SHAPE,
                  "local_conv1x1_bwd: bad shape (F <= 8 supported)");
    return launch1d("local_conv1x1_bwd", local_conv_bwd_kernel, (int64_t)H * W * Cin, as_stream(stream), x,
                    x_ld, dy, dy_ld, w, dx, dx_ld, dw, db, N, (int64_t)H * W, Cin, F);
}

int dl4ds_convlstm_gates_fwd(const float* z, const float* c_prev, float* c, float* h, int h_ld,
                             float* gates, int64_t n_pix, int F, void* stream) {
    DL4DS_REQUIRE(z && c && h && gates, DL4DS_E_BADARG, "convlstm_gates_fwd: null pointer");
    return launch1d("convlstm_gates_fwd", convlstm_gates_fwd_kernel, n_pix * F, as_stream(stream), z, c_prev,
                    c, h, h_ld, gates, n_pix, F);
}

int dl4ds_convlstm_gates_bwd(const float* gates, const float* c_prev, const float* c,
                             const float* dh, int dh_ld, const float* dc_next,
                             float* dz, float* dc_prev, int64_t n_pix, int F, void* stream) {
    DL4DS_REQUIRE(gates && c && dh && dz && dc_prev, DL4DS_E_BADARG, "convlstm_gates_bwd: null pointer");
    return launch1d("convlstm_gates_bwd", convlstm_gates_bwd_kernel, n_pix * F, as_stream(stream), gates,
                    c_prev, c, dh, dh_ld, dc_next, dz, dc_prev, n_pix, F);
}

}  // extern "C"
