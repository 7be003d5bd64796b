// Dropout variants of the conv blocks -- blocks.py:659-706 (`get_dropout_layer`): Dropout, GaussianDropout,
// SpatialDropout2D and their always-on Monte-Carlo twins.  TensorFlow's random streams cannot be reproduced, so
// the masks come from an own counter-based generator: Philox4x32-10 keyed by a seed kept in DEVICE memory, with the
// counter (element index, layer id, step).  A mask is a pure function of those, so nothing is stored: the backward
// pass regenerates it (dx = dy * mask, the same kernel), and a captured CUDA graph draws fresh masks at every
// replay because `dl4ds_rng_advance` (a kernel inside the graph) bumps the step.
#include <algorithm>

#include "common.cuh"

namespace dl4ds {
namespace {

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }   // [0, 1)

// variant 0: Dropout -- keep where u >= rate, scale 1/(1-rate)
//         1: GaussianDropout -- multiply by N(1, sqrt(rate / (1 - rate)))
//         2: SpatialDropout2D / 3D -- as 0 with one draw per (sample, channel); sample = (pixel / pix_per_sample) %
//            n_samples, so time-major frame stacks (T*B, H, W, C) share the draw over T as SpatialDropout3D does
// The mask of logical element idx (= flat NHWC index, or sample*C + c for the spatial variant) is a pure function:
//   variants 0 / 2: word (idx & 3) of Philox(counter idx >> 2);  variant 1: words 2*(idx & 1), +1 of Philox(idx >> 1).
// One thread owns four consecutive elements; with C % 4 == 0 they share one Philox call (two for the gaussian).
struct DropArgs {
    const float* x; float* y;
    int x_ld, y_ld;
    int64_t n_pix, pix_per_sample;
    int n_samples, C;
    float rate;
    int variant, layer_id, vec;
};

__device__ __forceinline__ float gauss_from(uint32_t a, uint32_t b, float sd) {
    const float u1 = 1.0f - u01(a), u2 = u01(b);                       // u1 in (0, 1]
    return 1.0f + sd * sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);    // Box-Muller
}

__global__ void __launch_bounds__(256) dropout_kernel(DropArgs a, const unsigned long long* __restrict__ state) {
    const unsigned long long seed = state[0], step = state[1];
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32);
    const float scale = 1.0f / (1.0f - a.rate);
    const float sd = sqrtf(a.rate / (1.0f - a.rate));
    const int64_t n = a.n_pix * a.C;
    const int64_t groups = (n + 3) >> 2;
    const bool shared_call = (a.C & 3) == 0;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = g << 2;
        float m[4];
        int64_t p0 = i0 / a.C;
        int c0 = (int)(i0 - p0 * a.C);
        if (shared_call) {
            const unsigned long long smp = (unsigned long long)((p0 / a.pix_per_sample) % a.n_samples);
            const unsigned long long idx0 = a.variant == 2 ? smp * a.C + c0 : (a.variant == 3 ? smp : (unsigned long long)i0);
            if (a.variant == 1) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const unsigned long long q = (idx0 >> 1) + h;
                    uint32_t ctr[4] = {(uint32_t)q, (uint32_t)(q >> 32), (uint32_t)a.layer_id, (uint32_t)step};
                    philox4x32_10(ctr, k0, k1);
                    m[2 * h] = gauss_from(ctr[0], ctr[1], sd);
                    m[2 * h + 1] = gauss_from(ctr[2], ctr[3], sd);
                }
            } else {
                const unsigned long long q = idx0 >> 2;
                uint32_t ctr[4] = {(uint32_t)q, (uint32_t)(q >> 32), (uint32_t)a.layer_id, (uint32_t)step};
                philox4x32_10(ctr, k0, k1);
#pragma unroll
                for (int k = 0; k < 4; ++k) m[k] = u01(ctr[k]) >= a.rate ? scale : 0.0f;
                if (a.variant == 3) {       // DropPath: one draw per sample, shared by every element of the sample
                    const int w = (int)(idx0 & 3);
                    m[0] = m[1] = m[2] = m[3] = w == 0 ? m[0] : w == 1 ? m[1] : w == 2 ? m[2] : m[3];
                }
            }
            if (a.vec) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(a.x + p0 * a.x_ld + c0));
                *reinterpret_cast<float4*>(a.y + p0 * a.y_ld + c0) = make_float4(v.x * m[0], v.y * m[1], v.z * m[2], v.w * m[3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) a.y[p0 * a.y_ld + c0 + k] = __ldg(a.x + p0 * a.x_ld + c0 + k) * m[k];
            }
            continue;
        }
        for (int k = 0; k < 4 && i0 + k < n; ++k) {     // any C: one call per element, same mask function
            const int64_t i = i0 + k, p = i / a.C;
            const int c = (int)(i - p * a.C);
            const unsigned long long smp = (unsigned long long)((p / a.pix_per_sample) % a.n_samples);
            const unsigned long long idx = a.variant == 2 ? smp * a.C + c : (a.variant == 3 ? smp : (unsigned long long)i);
            const unsigned long long q = a.variant == 1 ? idx >> 1 : idx >> 2;
            uint32_t ctr[4] = {(uint32_t)q, (uint32_t)(q >> 32), (uint32_t)a.layer_id, (uint32_t)step};
            philox4x32_10(ctr, k0, k1);
            float mk;
            if (a.variant == 1) {
                const int w = 2 * (int)(idx & 1);
                mk = gauss_from(w ? ctr[2] : ctr[0], w ? ctr[3] : ctr[1], sd);
            } else {
                const int w = (int)(idx & 3);
                mk = u01(w == 0 ? ctr[0] : w == 1 ? ctr[1] : w == 2 ? ctr[2] : ctr[3]) >= a.rate ? scale : 0.0f;
            }
            a.y[p * a.y_ld + c] = __ldg(a.x + p * a.x_ld + c) * mk;
        }
    }
}

__global__ void rng_advance_kernel(unsigned long long* state) {
    if (threadIdx.x == 0 && blockIdx.x == 0) state[1] += 1ull;
}

}  // namespace
}  // namespace dl4ds

using namespace dl4ds;

extern "C" {

int dl4ds_dropout(const float* x, int x_ld, float* y, int y_ld, int64_t n_pix, int64_t pix_per_sample, int n_samples,
                  int C, float rate, int variant, const uint64_t* rng_state, int layer_id, void* stream) {
    DL4DS_REQUIRE(x && y && rng_state, DL4DS_E_BADARG, "dropout: null pointer");
    DL4DS_REQUIRE(n_pix > 0 && C > 0 && pix_per_sample > 0 && n_pix % pix_per_sample == 0 && n_samples > 0,
                  DL4DS_E_SHAPE, "dropout: bad shape");
    DL4DS_REQUIRE(rate > 0.0f && rate < 1.0f, DL4DS_E_BADARG, "dropout: rate must be in (0, 1)");
    DL4DS_REQUIRE(variant >= 0 && variant <= 3, DL4DS_E_BADARG, "dropout: variant must be 0, 1, 2 or 3");
    const int64_t n = n_pix * C;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, 256 * 4 * 2), 16 * kNumSMs));
    DropArgs a{x, y, x_ld, y_ld, n_pix, pix_per_sample, n_samples, C, rate, variant, layer_id, 0};
    a.vec = (C % 4 == 0 && x_ld % 4 == 0 && y_ld % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) ? 1 : 0;
    dropout_kernel<<<grid, 256, 0, as_stream(stream)>>>(a, reinterpret_cast<const unsigned long long*>(rng_state));
    return check_launch("dropout");
}

int dl4ds_rng_advance(uint64_t* rng_state, void* stream) {
    DL4DS_REQUIRE(rng_state, DL4DS_E_BADARG, "rng_advance: null pointer");
    rng_advance_kernel<<<1, 32, 0, as_stream(stream)>>>(reinterpret_cast<unsigned long long*>(rng_state));
    return check_launch("rng_advance");
}

}  // extern "C"
