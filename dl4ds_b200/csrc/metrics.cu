// Verification metrics as device reductions -- dl4ds/metrics.py:15-97 (compute_rmse / compute_correlation over
// 'time' and 'space') and :166-186 (PSNR, MAE, dynamic range) of compute_metrics.  One pass per direction over the
// (N, P) fp32 pair (N = samples / time steps, P = H*W*C grid values per sample) produces the seven raw moments every
// one of those metrics is a function of, accumulated in fp64:
//     sum (y - yh)^2, sum |y - yh|, sum y, sum yh, sum y^2, sum yh^2, sum y*yh
//   * per sample ("over space": one block per sample, also min / max of y and yh for the dynamic range), and
//   * per grid value ("over time": one thread per grid value walking the N samples; consecutive threads read
//     consecutive addresses).
// The reference evaluates the same sums point by point on the host (joblib over grid points, scipy / sklearn
// per call); the closed forms (MSE, RMSE, Pearson r, mean bias, PSNR) are applied on the host from these moments.
#include "common.cuh"

namespace dl4ds {
namespace {

constexpr int kM = 7;      // moments per row
constexpr int kPair = 11;  // per-sample record: 7 moments + min y, max y, min yh, max yh

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256) metrics_pair_kernel(const float* __restrict__ y, const float* __restrict__ yh,
                                                           int64_t P, double* __restrict__ out) {
    __shared__ double red[8][kPair];
    const int64_t base = (int64_t)blockIdx.x * P;
    double m[kM] = {0, 0, 0, 0, 0, 0, 0};
    float mn = INFINITY, mx = -INFINITY, mnh = INFINITY, mxh = -INFINITY;
    for (int64_t i = threadIdx.x; i < P; i += 256) {
        const float a = __ldg(y + base + i), b = __ldg(yh + base + i);
        const double da = a, db = b, d = da - db;
        m[0] += d * d; m[1] += fabs(d); m[2] += da; m[3] += db; m[4] += da * da; m[5] += db * db; m[6] += da * db;
        mn = fminf(mn, a); mx = fmaxf(mx, a); mnh = fminf(mnh, b); mxh = fmaxf(mxh, b);
    }
#pragma unroll
    for (int k = 0; k < kM; ++k) m[k] = warp_sum_d(m[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mnh = fminf(mnh, __shfl_xor_sync(0xffffffffu, mnh, o)); mxh = fmaxf(mxh, __shfl_xor_sync(0xffffffffu, mxh, o));
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < kM; ++k) red[w][k] = m[k];
        red[w][7] = mn; red[w][8] = mx; red[w][9] = mnh; red[w][10] = mxh;
    }
    __syncthreads();
    if (threadIdx.x < kPair) {
        const int k = threadIdx.x;
        double v = red[0][k];
        for (int j = 1; j < 8; ++j) {
            const double u = red[j][k];
            if (k < kM) v += u;
            else if (k == 7 || k == 9) v = fmin(v, u);
            else v = fmax(v, u);
        }
        out[(int64_t)blockIdx.x * kPair + k] = v;
    }
}

__global__ void metrics_point_kernel(const float* __restrict__ y, const float* __restrict__ yh, int N, int64_t P,
                                     double* __restrict__ out) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        double m[kM] = {0, 0, 0, 0, 0, 0, 0};
        for (int n = 0; n < N; ++n) {
            const double da = __ldg(y + (int64_t)n * P + p), db = __ldg(yh + (int64_t)n * P + p), d = da - db;
            m[0] += d * d; m[1] += fabs(d); m[2] += da; m[3] += db; m[4] += da * da; m[5] += db * db; m[6] += da * db;
        }
#pragma unroll
        for (int k = 0; k < kM; ++k) out[p * kM + k] = m[k];
    }
}

}  // namespace
}  // namespace dl4ds

using namespace dl4ds;

extern "C" {

int dl4ds_metrics_moments(const float* y, const float* y_hat, int N, int64_t P, double* pair_out, double* point_out,
                          void* stream) {
    DL4DS_REQUIRE(y && y_hat && (pair_out || point_out), DL4DS_E_BADARG, "metrics_moments: null pointer");
    DL4DS_REQUIRE(N > 0 && P > 0, DL4DS_E_SHAPE, "metrics_moments: bad shape");
    cudaStream_t st = as_stream(stream);
    if (pair_out) metrics_pair_kernel<<<N, 256, 0, st>>>(y, y_hat, P, pair_out);
    if (point_out) {
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((P + 255) / 256, 16 * kNumSMs));
        metrics_point_kernel<<<grid, 256, 0, st>>>(y, y_hat, N, P, point_out);
    }
    return check_launch("metrics_moments");
}

}  // extern "C"
