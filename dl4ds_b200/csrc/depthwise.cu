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

// ---------------------------------------------------------------------------------------------------------
// 7x7 (the ConvNeXt size), tiled: a block owns CG channels x (TH x TW) pixels; the (TH+6) x (TW+6) x CG input tile
// sits in shared memory (row pitch padded so that the 32/CG pixel rows of a warp hit different banks); thread
// (channel, row) walks its row with the seven taps of one kernel row in registers: 7*TW FMAs per TW+6 loads.
// ---------------------------------------------------------------------------------------------------------
constexpr int kTW = 16;
template <int CG> struct Tile {
    static constexpr int TH = kThreads / CG;
    static constexpr int IW = kTW + 6, IH = TH + 6;
    static constexpr int PITCH = IW * CG + ((CG - (IW * CG) % 32 + 32) % 32);     // == CG (mod 32)
    static constexpr int FLOATS = IH * PITCH;
};

template <int CG>
__device__ __forceinline__ void load_tile(float* __restrict__ sm, const float* __restrict__ x, int x_ld, int n, int H,
                                          int W, int C, int h0, int w0, int c0) {
    using T = Tile<CG>;
    for (int i = threadIdx.x; i < T::IH * T::IW * CG; i += kThreads) {
        const int c = i % CG, q = (i / CG) % T::IW, r = i / (CG * T::IW);
        const int hh = h0 - 3 + r, ww = w0 - 3 + q;
        float v = 0.0f;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W && c0 + c < C)
            v = __ldg(x + (((int64_t)n * H + hh) * W + ww) * x_ld + c0 + c);
        sm[r * T::PITCH + q * CG + c] = v;
    }
}

template <int CG>
__global__ void __launch_bounds__(kThreads) depthwise7_fwd_kernel(const float* __restrict__ x, int x_ld,
                                                                  const float* __restrict__ wt,
                                                                  const float* __restrict__ bias,
                                                                  float* __restrict__ y, int y_ld, int N, int H, int W,
                                                                  int C, int flip, int accumulate, int tiles_w,
                                                                  int tiles_h) {
    using T = Tile<CG>;
    __shared__ float sm[T::FLOATS];
    const int c = threadIdx.x % CG, r = threadIdx.x / CG;
    int b = blockIdx.x;
    const int tw = b % tiles_w; b /= tiles_w;
    const int th = b % tiles_h; b /= tiles_h;
    const int n = b % N;
    const int c0 = (b / N) * CG;
    const int h0 = th * T::TH, w0 = tw * kTW;
    load_tile<CG>(sm, x, x_ld, n, H, W, C, h0, w0, c0);
    __syncthreads();
    const int cc = c0 + c, hh = h0 + r;
    if (cc >= C || hh >= H) return;
    float acc[kTW];
    const float b0 = bias ? __ldg(bias + cc) : 0.0f;
#pragma unroll
    for (int q = 0; q < kTW; ++q) acc[q] = b0;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        float wr[7], xr[T::IW];
#pragma unroll
        for (int j = 0; j < 7; ++j) wr[j] = __ldg(wt + (int64_t)(flip ? (6 - i) * 7 + (6 - j) : i * 7 + j) * C + cc);
#pragma unroll
        for (int q = 0; q < T::IW; ++q) xr[q] = sm[(r + i) * T::PITCH + q * CG + c];
#pragma unroll
        for (int q = 0; q < kTW; ++q)
#pragma unroll
            for (int j = 0; j < 7; ++j) acc[q] = fmaf(wr[j], xr[q + j], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < kTW; ++q) {
        if (w0 + q < W) {
            float* o = y + (((int64_t)n * H + hh) * W + w0 + q) * y_ld + cc;
            *o = accumulate ? *o + acc[q] : acc[q];
        }
    }
}

// weight gradient: blocks of one channel group walk over tiles, 49 accumulators per thread, one flush at the end
template <int CG>
__global__ void __launch_bounds__(kThreads) depthwise7_wgrad_kernel(const float* __restrict__ x, int x_ld,
                                                                    const float* __restrict__ dy, int dy_ld,
                                                                    float* __restrict__ dw, int N, int H, int W, int C,
                                                                    int tiles_w, int tiles_h, int workers) {
    using T = Tile<CG>;
    __shared__ float sm[T::FLOATS];
    __shared__ float red[49 * CG];
    const int c = threadIdx.x % CG, r = threadIdx.x / CG;
    const int c0 = (blockIdx.x / workers) * CG, worker = blockIdx.x % workers;
    const int cc = c0 + c;
    float acc[7][7];
#pragma unroll
    for (int i = 0; i < 7; ++i)
#pragma unroll
        for (int j = 0; j < 7; ++j) acc[i][j] = 0.0f;
    const int n_tiles = N * tiles_h * tiles_w;
    for (int t = worker; t < n_tiles; t += workers) {
        const int tw = t % tiles_w, th = (t / tiles_w) % tiles_h, n = t / (tiles_w * tiles_h);
        const int h0 = th * T::TH, w0 = tw * kTW;
        __syncthreads();
        load_tile<CG>(sm, x, x_ld, n, H, W, C, h0, w0, c0);
        __syncthreads();
        const int hh = h0 + r;
        float g[kTW];
#pragma unroll
        for (int q = 0; q < kTW; ++q)
            g[q] = (cc < C && hh < H && w0 + q < W) ? __ldg(dy + (((int64_t)n * H + hh) * W + w0 + q) * dy_ld + cc) : 0.0f;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            float xr[T::IW];
#pragma unroll
            for (int q = 0; q < T::IW; ++q) xr[q] = sm[(r + i) * T::PITCH + q * CG + c];
#pragma unroll
            for (int q = 0; q < kTW; ++q)
#pragma unroll
                for (int j = 0; j < 7; ++j) acc[i][j] = fmaf(g[q], xr[q + j], acc[i][j]);
        }
    }
    for (int i = threadIdx.x; i < 49 * CG; i += kThreads) red[i] = 0.0f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 7; ++i)
#pragma unroll
        for (int j = 0; j < 7; ++j) atomicAdd(red + (i * 7 + j) * CG + c, acc[i][j]);
    __syncthreads();
    for (int i = threadIdx.x; i < 49 * CG; i += kThreads) {
        const int tap = i / CG, ch = c0 + i % CG;
        if (ch < C) atomicAdd(dw + (int64_t)tap * C + ch, red[i]);
    }
}

template <int CG>
void launch_fwd7(const float* x, int x_ld, const float* w, const float* bias, float* y, int y_ld, int N, int H, int W,
                 int C, int flip, int accumulate, cudaStream_t st) {
    const int tiles_w = (int)cdiv(W, kTW), tiles_h = (int)cdiv(H, Tile<CG>::TH), groups = (int)cdiv(C, CG);
    depthwise7_fwd_kernel<CG><<<(unsigned)((int64_t)tiles_w * tiles_h * N * groups), kThreads, 0, st>>>(
        x, x_ld, w, bias, y, y_ld, N, H, W, C, flip, accumulate, tiles_w, tiles_h);
}

template <int CG>
void launch_wgrad7(const float* x, int x_ld, const float* dy, int dy_ld, float* dw, int N, int H, int W, int C,
                   cudaStream_t st) {
    const int tiles_w = (int)cdiv(W, kTW), tiles_h = (int)cdiv(H, Tile<CG>::TH), groups = (int)cdiv(C, CG);
    const int n_tiles = N * tiles_h * tiles_w;
    const int workers = std::max(1, std::min(n_tiles, (2 * kNumSMs + groups - 1) / groups));
    depthwise7_wgrad_kernel<CG><<<(unsigned)(groups * workers), kThreads, 0, st>>>(x, x_ld, dy, dy_ld, dw, N, H, W, C,
                                                                                   tiles_w, tiles_h, workers);
}

// channel-group width: 32 lanes when they are all busy, else the widest of 16 / 8 that wastes the fewest lanes
int pick_cg(int C) {
    if (C % 32 == 0) return 32;
    if (C % 16 == 0) return 16;
    if (C <= 8 || C % 8 == 0) return 8;
    return C > 16 ? 32 : 16;
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
    if (k == 7) {
        cudaStream_t st = as_stream(stream);
        switch (pick_cg(C)) {
            case 32: launch_fwd7<32>(x, x_ld, w, bias, y, y_ld, N, H, W, C, flip, accumulate, st); break;
            case 16: launch_fwd7<16>(x, x_ld, w, bias, y, y_ld, N, H, W, C, flip, accumulate, st); break;
            default: launch_fwd7<8>(x, x_ld, w, bias, y, y_ld, N, H, W, C, flip, accumulate, st); break;
        }
        return check_launch("depthwise_conv_fwd");
    }
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
    if (k == 7) {
        cudaStream_t st = as_stream(stream);
        switch (pick_cg(C)) {
            case 32: launch_wgrad7<32>(x, x_ld, dy, dy_ld, dw, N, H, W, C, st); break;
            case 16: launch_wgrad7<16>(x, x_ld, dy, dy_ld, dw, N, H, W, C, st); break;
            default: launch_wgrad7<8>(x, x_ld, dy, dy_ld, dw, N, H, W, C, st); break;
        }
        return check_launch("depthwise_conv_wgrad");
    }
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
