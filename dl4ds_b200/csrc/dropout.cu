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
__global__ void __launch_bounds__(256) dropout_kernel(const float* __restrict__ x, int x_ld, float* __restrict__ y,
                                                      int y_ld, int64_t n_pix, int64_t pix_per_sample, int n_samples,
                                                      int C, float rate, int variant,
                                                      const unsigned long long* __restrict__ state, int layer_id) {
    const unsigned long long seed = state[0], step = state[1];
    const float scale = 1.0f / (1.0f - rate);
    const float sd = sqrtf(rate / (1.0f - rate));
    const int64_t n = n_pix * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / C;
        const int c = (int)(i - p * C);
        const unsigned long long idx = variant == 2 ? (unsigned long long)((p / pix_per_sample) % n_samples) * C + c
                                                    : (unsigned long long)i;
        uint32_t ctr[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)layer_id, (uint32_t)step};
        philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32));
        float m;
        if (variant == 1) {
            const float u1 = 1.0f - u01(ctr[0]), u2 = u01(ctr[1]);          // u1 in (0, 1]
            m = 1.0f + sd * sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);   // Box-Muller
        } else {
            m = u01(ctr[0]) >= rate ? scale : 0.0f;
        }
        y[p * y_ld + c] = __ldg(x + p * x_ld + c) * m;
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
    DL4DS_REQUIRE(variant >= 0 && variant <= 2, DL4DS_E_BADARG, "dropout: variant must be 0, 1 or 2");
    const int64_t n = n_pix * C;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, 256 * 4), 8 * kNumSMs));
    dropout_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, x_ld, y, y_ld, n_pix, pix_per_sample, n_samples, C, rate,
                                                        variant,
                                                        reinterpret_cast<const unsigned long long*>(rng_state),
                                                        layer_id);
    return check_launch("dropout");
}

int dl4ds_rng_advance(uint64_t* rng_state, void* stream) {
    DL4DS_REQUIRE(rng_state, DL4DS_E_BADARG, "rng_advance: null pointer");
    rng_advance_kernel<<<1, 32, 0, as_stream(stream)>>>(reinterpret_cast<unsigned long long*>(rng_state));
    return check_launch("rng_advance");
}

}  // extern "C"
