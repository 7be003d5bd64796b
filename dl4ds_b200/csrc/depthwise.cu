// ConvNextBlock pieces that are not dense convolutions -- blocks.py:131-184:
//   DepthwiseConv2D(kernel_size=7, padding='same', depth_multiplier=1) forward / input gradient / weight gradient
//   and the exact (erf) GELU, the block's default activation.
// The pointwise Dense layers of the block run on the convolution kernels (1x1), LayerNormalization in norm.cu.
//
// Depthwise convolutions do k*k MACs per element: HBM/L1-bound CUDA-core work, no GEMM shape to give the tensor
// cores.  Mapping as in norm.cu: a group of G lanes owns one pixel, lane l the channels l, l+G, ..., so a warp
// touches 32 consecutive floats of the NHWC tensor per tap.
#include <algorithm>

#include "common.cuh"

namespace dl4ds {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxK = 7;

int group_width(int C) {
    int g = 1;
    while (g < C && g < 32) g <<= 1;
    return g;
}

// y[n,h,w,c] = bias[c] + sum_{i,j} wt[i][j][c] * x[n, h+i-r, w+j-r, c]   (flip: wt[k-1-i][k-1-j] -> input gradient)
__global__ void __launch_bounds__(kThreads) depthwise_fwd_kernel(const float* __restrict__ x, int x_ld,
                                                                 const float* __restrict__ wt,
                                                                 const float* __restrict__ bias,
                                                                 float* __restrict__ y, int y_ld, int N, int H, int W,
                                                                 int C, int k, int flip, int G, int accumulate) {
    const int lane = threadIdx.x % G, row = threadIdx.x / G, rows = kThreads / G;
    const int r = k / 2;
    const int64_t n_pix = (int64_t)N * H * W;
    for (int64_t p = (int64_t)blockIdx.x * rows + row; p < n_pix; p += (int64_t)gridDim.x * rows) {
        const int w0 = (int)(p % W), h0 = (int)((p / W) % H);
        const int64_t img = p / ((int64_t)H * W) * H * W;
        for (int c = lane; c < C; c += G) {
            float acc = bias ? __ldg(bias + c) : 0.0f;
            for (int i = 0; i < k; ++i) {
                const int hh = h0 + i - r;
                if (hh < 0 || hh >= H) continue;
                for (int j = 0; j < k; ++j) {
                    const int ww = w0 + j - r;
                    if (ww < 0 || ww >= W) continue;
                    const int t = flip ? ((k - 1 - i) * k + (k - 1 - j)) : (i * k + j);
                    acc = fmaf(__ldg(wt + (int64_t)t * C + c), __ldg(x + (img + (int64_t)hh * W + ww) * x_ld + c), acc);
                }
            }
            float* o = y + p * y_ld + c;
            *o = accumulate ? *o + acc : acc;
        }
    }
}

// dw[i][j][c] += sum_p dy[p,c] * x[p + (i-r, j-r), c]: 49 register accumulators per thread (one channel chunk at a time)
__global__ void __launch_bounds__(kThreads) depthwise_wgrad_kernel(const float* __restrict__ x, int x_ld,
                                                                   const float* __restrict__ dy, int dy_ld,
                                                                   float* __restrict__ dw, int N, int H, int W, int C,
                                                                   int k, int G) {
    __shared__ float sh[kMaxK * kMaxK * 32];
    const int lane = threadIdx.x % G, row = threadIdx.x / G, rows = kThreads / G;
    const int r = k / 2, kk = k * k;
    const int64_t n_pix = (int64_t)N * H * W;
    for (int c0 = 0; c0 < C; c0 += G) {
        const int c = c0 + lane;
        float acc[kMaxK * kMaxK];
#pragma unroll
        for (int t = 0; t < kMaxK * kMaxK; ++t) acc[t] = 0.0f;
        if (c < C) {
            for (int64_t p = (int64_t)blockIdx.x * rows + row; p < n_pix; p += (int64_t)gridDim.x * rows) {
                const int w0 = (int)(p % W), h0 = (int)((p / W) % H);
                const int64_t img = p / ((int64_t)H * W) * H * W;
                const float g = __ldg(dy + p * dy_ld + c);
#pragma unroll
                for (int i = 0; i < kMaxK; ++i) {
                    const int hh = h0 + i - r;
                    if (i >= k || hh < 0 || hh >= H) continue;
#pragma unroll
                    for (int j = 0; j < kMaxK; ++j) {
                        const int ww = w0 + j - r;
                        if (j >= k || ww < 0 || ww >= W) continue;
                        acc[i * kMaxK + j] = fmaf(g, __ldg(x + (img + (int64_t)hh * W + ww) * x_ld + c), acc[i * kMaxK + j]);
                    }
                }
            }
        }
        __syncthreads();
        for (int t = threadIdx.x; t < kk * G; t += kThreads) sh[t] = 0.0f;
        __syncthreads();
        if (c < C) {
#pragma unroll
            for (int i = 0; i < kMaxK; ++i)
#pragma unroll
                for (int j = 0; j < kMaxK; ++j)
                    if (i < k && j < k) atomicAdd(sh + (i * k + j) * G + lane, acc[i * kMaxK + j]);
        }
        __syncthreads();
        for (int t = threadIdx.x; t < kk * G; t += kThreads) {
            const int tap = t / G, cc = c0 + t % G;
            if (cc < C) atomicAdd(dw + (int64_t)tap * C + cc, sh[t]);
        }
    }
}

// exact GELU (Keras `gelu`, approximate=False): x * Phi(x)
__global__ void gelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = __ldg(x + i);
        y[i] = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    }
}

// dx = dy * (Phi(x) + x * phi(x))
__global__ void gelu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                                int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = __ldg(x + i);
        const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752f));
        const float pdf = 0.3989422804014327f * expf(-0.5f * v * v);
        dx[i] = __ldg(dy + i) * (cdf + v * pdf);
    }
}

int grid_rows(int64_t n_pix, int G, int max_blocks) {
    const int rows = kThreads / G;
    return (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(n_pix, rows), max_blocks));
}

}  // namespace
}  // namespace dl4ds

using namespace dl4ds;

extern "C" {

int dl4ds_depthwise_conv_fwd(const float* x, int x_ld, const float* w, const float* bias, float* y, int y_ld,
                             int N, int H, int W, int C, int k, int flip, int accumulate, void* stream) {
    DL4DS_REQUIRE(x && w && y && x != y, DL4DS_E_BADARG, "depthwise_conv_fwd: null or aliased pointer");
    DL4DS_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, DL4DS_E_SHAPE, "depthwise_conv_fwd: bad shape");
    DL4DS_REQUIRE(k >= 1 && k <= kMaxK && (k & 1), DL4DS_E_UNSUPPORTED, "depthwise_conv_fwd: k must be odd, <= %d",
                  kMaxK);
    const int G = group_width(C);
    const int64_t n_pix = (int64_t)N * H * W;
    depthwise_fwd_kernel<<<grid_rows(n_pix, G, 16 * kNumSMs), kThreads, 0, as_stream(stream)>>>(
        x, x_ld, w, bias, y, y_ld, N, H, W, C, k, flip, G, accumulate);
    return check_launch("depthwise_conv_fwd");
}

int dl4ds_depthwise_conv_wgrad(const float* x, int x_ld, const float* dy, int dy_ld, float* dw, int N, int H, int W,
                               int C, int k, void* stream) {
    DL4DS_REQUIRE(x && dy && dw, DL4DS_E_BADARG, "depthwise_conv_wgrad: null pointer");
    DL4DS_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, DL4DS_E_SHAPE, "depthwise_conv_wgrad: bad shape");
    DL4DS_REQUIRE(k >= 1 && k <= kMaxK && (k & 1), DL4DS_E_UNSUPPORTED, "depthwise_conv_wgrad: k must be odd, <= %d",
                  kMaxK);
    const int G = group_width(C);
    const int64_t n_pix = (int64_t)N * H * W;
    depthwise_wgrad_kernel<<<grid_rows(n_pix, G, 2 * kNumSMs), kThreads, 0, as_stream(stream)>>>(
        x, x_ld, dy, dy_ld, dw, N, H, W, C, k, G);
    return check_launch("depthwise_conv_wgrad");
}

int dl4ds_gelu_fwd(const float* x, float* y, int64_t n, void* stream) {
    DL4DS_REQUIRE(x && y && n > 0, DL4DS_E_BADARG, "gelu_fwd: null pointer or n <= 0");
    const int grid = (int)std::min<int64_t>(cdiv(n, 256 * 4), 8 * kNumSMs);
    gelu_fwd_kernel<<<std::max(grid, 1), 256, 0, as_stream(stream)>>>(x, y, n);
    return check_launch("gelu_fwd");
}

int dl4ds_gelu_bwd(const float* x, const float* dy, float* dx, int64_t n, void* stream) {
    DL4DS_REQUIRE(x && dy && dx && n > 0, DL4DS_E_BADARG, "gelu_bwd: null pointer or n <= 0");
    const int grid = (int)std::min<int64_t>(cdiv(n, 256 * 4), 8 * kNumSMs);
    gelu_bwd_kernel<<<std::max(grid, 1), 256, 0, as_stream(stream)>>>(x, dy, dx, n);
    return check_launch("gelu_bwd");
}

}  // extern "C"
