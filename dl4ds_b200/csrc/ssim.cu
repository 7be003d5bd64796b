// SSIM-family losses (losses.py:23-147): DSSIM and multi-scale DSSIM with their exact gradients.
//
// The reference calls tf.image.ssim / tf.image.ssim_multiscale (11x11 gaussian, sigma 1.5, k1 0.01, k2 0.03,
// VALID depthwise filtering, max_val = dynamic range of both tensors, each tensor shifted by its own minimum
// when that is negative -- losses.py:44-57,118-131).  Here:
//
//   range kernels     min/max (+ arg) of y_pred, y_true -> shifts, L, C1 = (k1 L)^2, C2 = (k2 L)^2 in device memory
//   ssim_maps_kernel  per 32x32 tile of window positions: separable gaussian moments of x, y, xy, x^2+y^2 in shared
//                     memory, SSIM (or cs only, scales before the last of MS-SSIM), and the three partial-derivative
//                     maps dS/d(mean x), dS/d(E x^2), dS/d(E xy); per-plane sums of S, dS/dC1, dS/dC2
//   ssim_combine      per plane: relu, weighted geometric mean over scales, loss, per-scale gradient coefficients,
//                     gradient with respect to the dynamic range
//   ssim_bwd_kernel   d loss / d x(q) = coef * [ G^T(D1) + 2 x(q) G^T(D2) + y(q) G^T(D3) ] (+ 1/4 of the coarser
//                     scale's gradient: the 2x2 average pooling between scales), G^T = transposed gaussian filtering,
//                     again separable in shared memory
//   ssim_fixup        the gradient through max / min (dynamic range) and through the shift lands on the arg-max /
//                     arg-min elements of y_pred (tf.reduce_max / reduce_min / maximum / minimum gradients)
//
// Everything is HBM-bound streaming work: 8 B/pixel in, 12 B/window position out (maps), the reverse in the backward.
#include <algorithm>
#include <math.h>

#include "common.cuh"

namespace dl4ds {
namespace {

constexpr int kWin = 11;
constexpr int kHalo = kWin - 1;
constexpr int kTile = 32;
constexpr int kIn = kTile + kHalo;   // 42
constexpr int kMaxScales = 5;
constexpr int kRangeBlocks = 2 * kNumSMs;
constexpr float kK1 = 0.01f, kK2 = 0.03f;

struct Gauss { float w[kWin]; };

// stats slots
enum { S_SHIFT_P = 0, S_SHIFT_T, S_C1, S_C2, S_L, S_MAX_IN_PRED, S_MIN_IN_PRED, S_SHIFTED_P, S_DL, S_SUMDX,
       S_IMAX, S_IMAX_HI, S_IMIN, S_IMIN_HI, S_COUNT = 16 };

struct Partial { float maxp, minp, maxt, mint; long long imax, imin; };

__global__ void __launch_bounds__(256) range_partial_kernel(const float* __restrict__ yp,
                                                            const float* __restrict__ yt, int64_t n,
                                                            Partial* __restrict__ part) {
    float maxp = -INFINITY, minp = INFINITY, maxt = -INFINITY, mint = INFINITY;
    long long imax = 0, imin = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float p = __ldg(yp + i), t = __ldg(yt + i);
        if (p > maxp) { maxp = p; imax = i; }      // ascending i per thread: first occurrence wins
        if (p < minp) { minp = p; imin = i; }
        maxt = fmaxf(maxt, t);
        mint = fminf(mint, t);
    }
    __shared__ Partial sh[256];
    sh[threadIdx.x] = Partial{maxp, minp, maxt, mint, imax, imin};
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            Partial a = sh[threadIdx.x];
            const Partial b = sh[threadIdx.x + s];
            if (b.maxp > a.maxp || (b.maxp == a.maxp && b.imax < a.imax)) { a.maxp = b.maxp; a.imax = b.imax; }
            if (b.minp < a.minp || (b.minp == a.minp && b.imin < a.imin)) { a.minp = b.minp; a.imin = b.imin; }
            a.maxt = fmaxf(a.maxt, b.maxt);
            a.mint = fminf(a.mint, b.mint);
            sh[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

// one block: reduce the partials, derive the scalars, zero the accumulators that follow the stats block
__global__ void __launch_bounds__(256) range_final_kernel(const Partial* __restrict__ part, int n_part,
                                                          float* __restrict__ stats, int n_zero) {
    __shared__ Partial sh[256];
    Partial a{-INFINITY, INFINITY, -INFINITY, INFINITY, 0, 0};
    for (int i = threadIdx.x; i < n_part; i += blockDim.x) {
        const Partial b = part[i];
        if (b.maxp > a.maxp || (b.maxp == a.maxp && b.imax < a.imax)) { a.maxp = b.maxp; a.imax = b.imax; }
        if (b.minp < a.minp || (b.minp == a.minp && b.imin < a.imin)) { a.minp = b.minp; a.imin = b.imin; }
        a.maxt = fmaxf(a.maxt, b.maxt);
        a.mint = fminf(a.mint, b.mint);
    }
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            Partial x = sh[threadIdx.x];
            const Partial b = sh[threadIdx.x + s];
            if (b.maxp > x.maxp || (b.maxp == x.maxp && b.imax < x.imax)) { x.maxp = b.maxp; x.imax = b.imax; }
            if (b.minp < x.minp || (b.minp == x.minp && b.imin < x.imin)) { x.minp = b.minp; x.imin = b.imin; }
            x.maxt = fmaxf(x.maxt, b.maxt);
            x.mint = fminf(x.mint, b.mint);
            sh[threadIdx.x] = x;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n_zero; i += blockDim.x) stats[S_COUNT + i] = 0.0f;
    if (threadIdx.x == 0) {
        const Partial r = sh[0];
        // losses.py:44-46: maxv = maximum(max(y_true), max(y_pred)); minv = minimum(min(y_true), min(y_pred))
        const float L = fmaxf(r.maxt, r.maxp) - fminf(r.mint, r.minp);
        stats[S_SHIFT_P] = r.minp < 0.0f ? r.minp : 0.0f;       // losses.py:51-54
        stats[S_SHIFT_T] = r.mint < 0.0f ? r.mint : 0.0f;       // losses.py:47-50
        stats[S_C1] = (kK1 * L) * (kK1 * L);
        stats[S_C2] = (kK2 * L) * (kK2 * L);
        stats[S_L] = L;
        stats[S_MAX_IN_PRED] = r.maxp > r.maxt ? 1.0f : 0.0f;   // tf.maximum(a, b) sends the gradient to a when a >= b
        stats[S_MIN_IN_PRED] = r.minp < r.mint ? 1.0f : 0.0f;   // tf.minimum(a, b): to a when a <= b
        stats[S_SHIFTED_P] = r.minp < 0.0f ? 1.0f : 0.0f;
        stats[S_DL] = 0.0f;
        stats[S_SUMDX] = 0.0f;
        reinterpret_cast<long long*>(stats + S_IMAX)[0] = r.imax;
        reinterpret_cast<long long*>(stats + S_IMIN)[0] = r.imin;
    }
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = v;
    __syncthreads();
    float r = 0.0f;
    if (threadIdx.x < 32) {
        r = threadIdx.x < 8 ? red[threadIdx.x] : 0.0f;
        r = warp_sum(r);
    }
    return r;   // valid on thread 0
}

// grid (tiles_x, tiles_y, B*C); block 256.  maps: [3][B][Ho][Wo][C]; plane_acc: [3][B*C] (sum S, sum dS/dC1, sum dS/dC2)
__global__ void __launch_bounds__(256) ssim_maps_kernel(const float* __restrict__ yp, const float* __restrict__ yt,
                                                        const float* __restrict__ stats, int H, int W, int C,
                                                        int cs_only, Gauss g, float* __restrict__ maps,
                                                        float* __restrict__ plane_acc, int n_planes) {
    __shared__ float sx[kIn][kIn + 1], sy[kIn][kIn + 1];
    __shared__ float hrow[4][kIn][kTile + 1];
    __shared__ float red[8];
    const int Ho = H - kHalo, Wo = W - kHalo;
    const int plane = blockIdx.z, b = plane / C, c = plane % C;
    const int ox0 = blockIdx.x * kTile, oy0 = blockIdx.y * kTile;
    const float shp = stats[S_SHIFT_P], sht = stats[S_SHIFT_T], C1 = stats[S_C1], C2 = stats[S_C2];
    for (int i = threadIdx.x; i < kIn * kIn; i += 256) {
        const int r = i / kIn, q = i % kIn;
        const int yy = oy0 + r, xx = ox0 + q;
        float vx = 0.0f, vy = 0.0f;
        if (yy < H && xx < W) {
            const int64_t o = (((int64_t)b * H + yy) * W + xx) * C + c;
            vx = __ldg(yp + o) - shp;
            vy = __ldg(yt + o) - sht;
        }
        sx[r][q] = vx;
        sy[r][q] = vy;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kIn * kTile; i += 256) {
        const int r = i / kTile, q = i % kTile;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
        for (int t = 0; t < kWin; ++t) {
            const float vx = sx[r][q + t], vy = sy[r][q + t], w = g.w[t];
            a0 = fmaf(w, vx, a0);
            a1 = fmaf(w, vy, a1);
            a2 = fmaf(w, vx * vy, a2);
            a3 = fmaf(w, fmaf(vx, vx, vy * vy), a3);
        }
        hrow[0][r][q] = a0; hrow[1][r][q] = a1; hrow[2][r][q] = a2; hrow[3][r][q] = a3;
    }
    __syncthreads();
    float sumS = 0.0f, sumC1 = 0.0f, sumC2 = 0.0f;
    const int64_t map_stride = (int64_t)n_planes * Ho * Wo;
    for (int i = threadIdx.x; i < kTile * kTile; i += 256) {
        const int r = i / kTile, q = i % kTile;
        const int oy = oy0 + r, ox = ox0 + q;
        if (oy >= Ho || ox >= Wo) continue;
        float mx = 0.0f, my = 0.0f, exy = 0.0f, e2 = 0.0f;
#pragma unroll
        for (int t = 0; t < kWin; ++t) {
            const float w = g.w[t];
            mx = fmaf(w, hrow[0][r + t][q], mx);
            my = fmaf(w, hrow[1][r + t][q], my);
            exy = fmaf(w, hrow[2][r + t][q], exy);
            e2 = fmaf(w, hrow[3][r + t][q], e2);
        }
        // _ssim_helper: luminance = (2 mx my + c1) / (mx^2 + my^2 + c1); cs = (2 E[xy] - 2 mx my + c2) / (E[x^2+y^2] - mx^2 - my^2 + c2)
        const float num0 = 2.0f * mx * my, den0 = mx * mx + my * my;
        const float B1 = den0 + C1, B2 = e2 - den0 + C2;
        const float iB1 = 1.0f / B1, iB2 = 1.0f / B2;
        float lum = (num0 + C1) * iB1;
        const float cs = (2.0f * exy - num0 + C2) * iB2;
        float dlum_dmx = (2.0f * my - lum * 2.0f * mx) * iB1;
        float dlum_dC1 = (1.0f - lum) * iB1;
        if (cs_only) { lum = 1.0f; dlum_dmx = 0.0f; dlum_dC1 = 0.0f; }
        const float dcs_dmx = (2.0f * mx * cs - 2.0f * my) * iB2;
        const float S = lum * cs;
        const float d1 = cs * dlum_dmx + lum * dcs_dmx;   // dS / d mean(x)
        const float d2 = -lum * cs * iB2;                 // dS / d E[x^2]
        const float d3 = lum * 2.0f * iB2;                // dS / d E[xy]
        sumS += S;
        sumC1 += cs * dlum_dC1;
        sumC2 += lum * (1.0f - cs) * iB2;
        const int64_t o = (((int64_t)b * Ho + oy) * Wo + ox) * C + c;
        maps[o] = d1;
        maps[o + map_stride] = d2;
        maps[o + 2 * map_stride] = d3;
    }
    const float t0 = block_sum_256(sumS, red);
    const float t1 = block_sum_256(sumC1, red);
    const float t2 = block_sum_256(sumC2, red);
    if (threadIdx.x == 0) {
        atomicAdd(plane_acc + plane, t0);
        atomicAdd(plane_acc + n_planes + plane, t1);
        atomicAdd(plane_acc + 2 * n_planes + plane, t2);
    }
}

// ssim_multiscale's downsampling: pad odd sizes by one (mode SYMMETRIC = repeat the edge), avg_pool 2x2 stride 2
__global__ void ssim_pool_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C) {
    const int Hc = (H + 1) / 2, Wc = (W + 1) / 2;
    const int64_t total = (int64_t)B * Hc * Wc * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t t = i / C;
        const int j = (int)(t % Wc); t /= Wc;
        const int r = (int)(t % Hc);
        const int b = (int)(t / Hc);
        const int r0 = 2 * r, r1 = min(2 * r + 1, H - 1), c0 = 2 * j, c1 = min(2 * j + 1, W - 1);
        const float* p = x + (int64_t)b * H * W * C + c;
        y[i] = 0.25f * ((__ldg(p + ((int64_t)r0 * W + c0) * C) + __ldg(p + ((int64_t)r0 * W + c1) * C)) +
                        (__ldg(p + ((int64_t)r1 * W + c0) * C) + __ldg(p + ((int64_t)r1 * W + c1) * C)));
    }
}

struct CombineArgs {
    int n_scales, n_planes, B, C;
    float pf[kMaxScales];
    float inv_count[kMaxScales];   // 1 / (Ho_j * Wo_j)
    float scale;
};

// plane_acc: [n_scales][3][n_planes]; coef: [n_scales][n_planes]
__global__ void ssim_combine_kernel(CombineArgs a, const float* __restrict__ plane_acc, float* __restrict__ stats,
                                    float* __restrict__ coef, float* __restrict__ loss_out) {
    const float L = stats[S_L];
    float loss = 0.0f, dL = 0.0f;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < a.n_planes; p += gridDim.x * blockDim.x) {
        float v[kMaxScales];
        float ms = 1.0f;
        for (int j = 0; j < a.n_scales; ++j) {
            const float raw = plane_acc[((int64_t)j * 3 + 0) * a.n_planes + p] * a.inv_count[j];
            v[j] = a.n_scales == 1 ? raw : fmaxf(raw, 0.0f);   // ssim_multiscale applies nn.relu per scale, ssim does not
            ms *= a.n_scales == 1 ? v[j] : powf(v[j], a.pf[j]);
        }
        // loss = scale * mean_b((1 - mean_c ms) / 2)
        const float base = -0.5f * a.scale / (float)a.n_planes;
        loss += a.scale * 0.5f * (1.0f - ms) / (float)a.n_planes;
        for (int j = 0; j < a.n_scales; ++j) {
            const float dms_dv = a.n_scales == 1 ? 1.0f : (v[j] > 0.0f ? a.pf[j] * ms / v[j] : 0.0f);
            const float cj = base * dms_dv * a.inv_count[j];
            coef[(int64_t)j * a.n_planes + p] = cj;
            dL += cj * (plane_acc[((int64_t)j * 3 + 1) * a.n_planes + p] * (2.0f * kK1 * kK1 * L) +
                        plane_acc[((int64_t)j * 3 + 2) * a.n_planes + p] * (2.0f * kK2 * kK2 * L));
        }
    }
    loss = warp_sum(loss);
    dL = warp_sum(dL);
    if (threadIdx.x % 32 == 0) {
        atomicAdd(loss_out, loss);
        atomicAdd(stats + S_DL, dL);
    }
}

// grid (tiles_x, tiles_y, B*C) over the H x W pixels.  dy (+)= coef[plane] * (...) + up/4
__global__ void __launch_bounds__(256) ssim_bwd_kernel(const float* __restrict__ yp, const float* __restrict__ yt,
                                                       float* __restrict__ stats, int H, int W, int C, Gauss g,
                                                       const float* __restrict__ maps, const float* __restrict__ coef,
                                                       const float* __restrict__ up, float* __restrict__ dy,
                                                       int accumulate, int n_planes, int want_sum) {
    __shared__ float sg[3][kIn][kIn + 1];
    __shared__ float hrow[3][kIn][kTile + 1];
    __shared__ float red[8];
    const int Ho = H - kHalo, Wo = W - kHalo;
    const int plane = blockIdx.z, b = plane / C, c = plane % C;
    const int qx0 = blockIdx.x * kTile, qy0 = blockIdx.y * kTile;
    const float shp = stats[S_SHIFT_P], sht = stats[S_SHIFT_T];
    const int64_t map_stride = (int64_t)n_planes * Ho * Wo;
    // tile index (r, j) <-> window position (qy0 - 10 + r, qx0 - 10 + j); outside the map = 0
    for (int i = threadIdx.x; i < kIn * kIn; i += 256) {
        const int r = i / kIn, j = i % kIn;
        const int py = qy0 - kHalo + r, px = qx0 - kHalo + j;
        float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f;
        if (py >= 0 && py < Ho && px >= 0 && px < Wo) {
            const int64_t o = (((int64_t)b * Ho + py) * Wo + px) * C + c;
            v0 = __ldg(maps + o);
            v1 = __ldg(maps + o + map_stride);
            v2 = __ldg(maps + o + 2 * map_stride);
        }
        sg[0][r][j] = v0; sg[1][r][j] = v1; sg[2][r][j] = v2;
    }
    __syncthreads();
    // G^T(D)(q) = sum_d w[d] D(q - d): along x first
    for (int i = threadIdx.x; i < kIn * kTile; i += 256) {
        const int r = i / kTile, q = i % kTile;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
#pragma unroll
        for (int d = 0; d < kWin; ++d) {
            const float w = g.w[d];
            a0 = fmaf(w, sg[0][r][q + kHalo - d], a0);
            a1 = fmaf(w, sg[1][r][q + kHalo - d], a1);
            a2 = fmaf(w, sg[2][r][q + kHalo - d], a2);
        }
        hrow[0][r][q] = a0; hrow[1][r][q] = a1; hrow[2][r][q] = a2;
    }
    __syncthreads();
    const float cf = __ldg(coef + plane);
    float sum = 0.0f;
    const int Hc = (H + 1) / 2, Wc = (W + 1) / 2;     // pooled size; an odd edge row / column was mirrored: it counts twice
    for (int i = threadIdx.x; i < kTile * kTile; i += 256) {
        const int r = i / kTile, q = i % kTile;
        const int yy = qy0 + r, xx = qx0 + q;
        if (yy >= H || xx >= W) continue;
        float t1 = 0.0f, t2 = 0.0f, t3 = 0.0f;
#pragma unroll
        for (int d = 0; d < kWin; ++d) {
            const float w = g.w[d];
            t1 = fmaf(w, hrow[0][r + kHalo - d][q], t1);
            t2 = fmaf(w, hrow[1][r + kHalo - d][q], t2);
            t3 = fmaf(w, hrow[2][r + kHalo - d][q], t3);
        }
        const int64_t o = (((int64_t)b * H + yy) * W + xx) * C + c;
        const float x = __ldg(yp + o) - shp, y = __ldg(yt + o) - sht;
        float gq = cf * (t1 + 2.0f * x * t2 + y * t3);
        if (up) {
            const float mult = (((H & 1) && yy == H - 1) ? 2.0f : 1.0f) * (((W & 1) && xx == W - 1) ? 2.0f : 1.0f);
            gq += 0.25f * mult * __ldg(up + (((int64_t)b * Hc + (yy >> 1)) * Wc + (xx >> 1)) * C + c);
        }
        sum += gq;
        dy[o] = accumulate ? dy[o] + gq : gq;
    }
    if (want_sum) {
        const float t = block_sum_256(sum, red);
        if (threadIdx.x == 0) atomicAdd(stats + S_SUMDX, t);
    }
}

__global__ void ssim_fixup_kernel(const float* __restrict__ stats, float* __restrict__ dy) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const long long imax = reinterpret_cast<const long long*>(stats + S_IMAX)[0];
    const long long imin = reinterpret_cast<const long long*>(stats + S_IMIN)[0];
    const float dL = stats[S_DL];
    float gmax = 0.0f, gmin = 0.0f;
    if (stats[S_MAX_IN_PRED] != 0.0f) gmax += dL;                 // L = maxv - minv
    if (stats[S_MIN_IN_PRED] != 0.0f) gmin -= dL;
    if (stats[S_SHIFTED_P] != 0.0f) gmin -= stats[S_SUMDX];       // x = y_pred - min(y_pred)
    if (imax == imin) {
        dy[imax] += gmax + gmin;
    } else {
        dy[imax] += gmax;
        dy[imin] += gmin;
    }
}

// tf.image.ssim(img1, img2, max_val) per image (metrics.py:172-176): no shift, C1 / C2 from the given dynamic range
__global__ void ssim_index_setup_kernel(float max_val, float* __restrict__ stats, float* __restrict__ plane_acc, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) plane_acc[i] = 0.0f;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        stats[S_SHIFT_P] = 0.0f;
        stats[S_SHIFT_T] = 0.0f;
        stats[S_C1] = (kK1 * max_val) * (kK1 * max_val);
        stats[S_C2] = (kK2 * max_val) * (kK2 * max_val);
        stats[S_L] = max_val;
    }
}

__global__ void ssim_index_final_kernel(const float* __restrict__ plane_acc, float inv_count, int B, int C,
                                        float* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float s = 0.0f;
    for (int c = 0; c < C; ++c) s += plane_acc[b * C + c] * inv_count;
    out[b] = s / (float)C;
}

struct Layout {
    int64_t stats, plane_acc, coef, partial, total;
    int64_t img_p[kMaxScales], img_t[kMaxScales], grad[kMaxScales], maps[kMaxScales];
    int Hs[kMaxScales], Ws[kMaxScales];
};

Layout make_layout(int B, int H, int W, int C, int nS) {
    Layout l{};
    auto align = [](int64_t v) { return (v + 3) & ~int64_t(3); };
    int64_t o = 0;
    l.stats = o; o += S_COUNT;
    l.plane_acc = o; o = align(o + (int64_t)nS * 3 * B * C);     // zeroed together with the stats accumulators
    l.coef = o; o = align(o + (int64_t)nS * B * C);
    l.partial = o; o = align(o + (int64_t)kRangeBlocks * (sizeof(Partial) / sizeof(float)));
    int h = H, w = W;
    for (int j = 0; j < nS; ++j) {
        l.Hs[j] = h; l.Ws[j] = w;
        if (j > 0) {
            l.img_p[j] = o; o = align(o + (int64_t)B * h * w * C);
            l.img_t[j] = o; o = align(o + (int64_t)B * h * w * C);
            l.grad[j] = o; o = align(o + (int64_t)B * h * w * C);
        }
        l.maps[j] = o; o = align(o + 3 * (int64_t)B * (h - kHalo) * (w - kHalo) * C);
        h = (h + 1) / 2; w = (w + 1) / 2;
    }
    l.total = o;
    return l;
}

int check_shape(int B, int H, int W, int C, int nS) {
    DL4DS_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, DL4DS_E_SHAPE, "ssim_loss: bad shape");
    DL4DS_REQUIRE(nS >= 1 && nS <= kMaxScales, DL4DS_E_BADARG, "ssim_loss: n_scales must be 1..%d", kMaxScales);
    DL4DS_REQUIRE((int64_t)B * C <= 65535, DL4DS_E_SHAPE, "ssim_loss: B*C > 65535");
    int h = H, w = W;
    for (int j = 0; j < nS; ++j) {
        // tf.image.ssim asserts every image is at least filter_size wide at every scale
        DL4DS_REQUIRE(h >= kWin && w >= kWin, DL4DS_E_SHAPE,
                      "ssim_loss: %dx%d at scale %d is smaller than the 11x11 window", h, w, j);
        h = (h + 1) / 2; w = (w + 1) / 2;       // odd sizes are SYMMETRIC-padded by one before the 2x2 pooling
    }
    return DL4DS_OK;
}

Gauss make_gauss() {
    // _fspecial_gauss: softmax over the 2-D grid of -(x^2 + y^2) / (2 sigma^2) = outer product of the normalised 1-D kernel
    Gauss g;
    double s = 0.0, e[kWin];
    for (int i = 0; i < kWin; ++i) {
        const double c = i - (kWin - 1) / 2.0;
        e[i] = exp(-0.5 * c * c / (1.5 * 1.5));
        s += e[i];
    }
    for (int i = 0; i < kWin; ++i) g.w[i] = (float)(e[i] / s);
    return g;
}

}  // namespace
}  // namespace dl4ds

using namespace dl4ds;

extern "C" {

int64_t dl4ds_ssim_loss_workspace_floats(int B, int H, int W, int C, int n_scales) {
    if (check_shape(B, H, W, C, n_scales) != DL4DS_OK) return -1;
    return make_layout(B, H, W, C, n_scales).total;
}

int dl4ds_ssim_loss(const float* y_pred, const float* y_true, int B, int H, int W, int C, int n_scales,
                    const float* power_factors, float scale, float* loss_out, float* dy, int accumulate,
                    float* ws, void* stream) {
    DL4DS_REQUIRE(y_pred && y_true && loss_out && ws, DL4DS_E_BADARG, "ssim_loss: null pointer");
    DL4DS_REQUIRE(n_scales == 1 || power_factors, DL4DS_E_BADARG, "ssim_loss: power_factors missing");
    DL4DS_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, DL4DS_E_BADARG, "ssim_loss: ws not 16-byte aligned");
    const int rc = check_shape(B, H, W, C, n_scales);
    if (rc != DL4DS_OK) return rc;
    cudaStream_t st = as_stream(stream);
    const Layout l = make_layout(B, H, W, C, n_scales);
    const Gauss g = make_gauss();
    const int n_planes = B * C;
    const int64_t n = (int64_t)B * H * W * C;
    float* stats = ws + l.stats;
    Partial* part = reinterpret_cast<Partial*>(ws + l.partial);

    const int rb = (int)std::min<int64_t>(kRangeBlocks, cdiv(n, 256));
    range_partial_kernel<<<rb, 256, 0, st>>>(y_pred, y_true, n, part);
    range_final_kernel<<<1, 256, 0, st>>>(part, rb, stats, (int)(l.coef - S_COUNT));

    CombineArgs ca{};
    ca.n_scales = n_scales; ca.n_planes = n_planes; ca.B = B; ca.C = C; ca.scale = scale;
    const float* xp = y_pred;
    const float* xt = y_true;
    for (int j = 0; j < n_scales; ++j) {
        const int h = l.Hs[j], w = l.Ws[j];
        if (j > 0) {   // ssim_multiscale: avg_pool(ksize 2, stride 2) between scales
            const int64_t np_ = (int64_t)B * h * w * C;
            const int pg = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(np_, 256), 8 * kNumSMs));
            ssim_pool_kernel<<<pg, 256, 0, st>>>(xp, ws + l.img_p[j], B, l.Hs[j - 1], l.Ws[j - 1], C);
            ssim_pool_kernel<<<pg, 256, 0, st>>>(xt, ws + l.img_t[j], B, l.Hs[j - 1], l.Ws[j - 1], C);
            xp = ws + l.img_p[j];
            xt = ws + l.img_t[j];
        }
        const int Ho = h - kHalo, Wo = w - kHalo;
        ca.pf[j] = n_scales == 1 ? 1.0f : power_factors[j];
        ca.inv_count[j] = 1.0f / ((float)Ho * (float)Wo);
        dim3 grid((unsigned)cdiv(Wo, kTile), (unsigned)cdiv(Ho, kTile), (unsigned)n_planes);
        ssim_maps_kernel<<<grid, 256, 0, st>>>(xp, xt, stats, h, w, C, j + 1 < n_scales ? 1 : 0, g, ws + l.maps[j],
                                               ws + l.plane_acc + (int64_t)j * 3 * n_planes, n_planes);
    }
    ssim_combine_kernel<<<(unsigned)cdiv(n_planes, 128), 128, 0, st>>>(ca, ws + l.plane_acc, stats, ws + l.coef,
                                                                       loss_out);
    if (dy) {
        for (int j = n_scales - 1; j >= 0; --j) {
            const int h = l.Hs[j], w = l.Ws[j];
            const float* ip = j == 0 ? y_pred : ws + l.img_p[j];
            const float* it = j == 0 ? y_true : ws + l.img_t[j];
            float* out = j == 0 ? dy : ws + l.grad[j];
            const float* up = j + 1 < n_scales ? ws + l.grad[j + 1] : nullptr;
            dim3 grid((unsigned)cdiv(w, kTile), (unsigned)cdiv(h, kTile), (unsigned)n_planes);
            ssim_bwd_kernel<<<grid, 256, 0, st>>>(ip, it, stats, h, w, C, g, ws + l.maps[j],
                                                  ws + l.coef + (int64_t)j * n_planes, up, out,
                                                  j == 0 ? accumulate : 0, n_planes, j == 0 ? 1 : 0);
        }
        ssim_fixup_kernel<<<1, 32, 0, st>>>(stats, dy);
    }
    return check_launch("ssim_loss");
}

int dl4ds_ssim_index(const float* img1, const float* img2, int B, int H, int W, int C, float max_val, float* out,
                     float* ws, void* stream) {
    DL4DS_REQUIRE(img1 && img2 && out && ws, DL4DS_E_BADARG, "ssim_index: null pointer");
    DL4DS_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, DL4DS_E_BADARG, "ssim_index: ws not 16-byte aligned");
    const int rc = check_shape(B, H, W, C, 1);
    if (rc != DL4DS_OK) return rc;
    cudaStream_t st = as_stream(stream);
    const Layout l = make_layout(B, H, W, C, 1);
    const Gauss g = make_gauss();
    const int n_planes = B * C;
    float* stats = ws + l.stats;
    ssim_index_setup_kernel<<<(unsigned)cdiv(3 * n_planes, 256), 256, 0, st>>>(max_val, stats, ws + l.plane_acc,
                                                                               3 * n_planes);
    const int Ho = H - kHalo, Wo = W - kHalo;
    dim3 grid((unsigned)cdiv(Wo, kTile), (unsigned)cdiv(Ho, kTile), (unsigned)n_planes);
    ssim_maps_kernel<<<grid, 256, 0, st>>>(img1, img2, stats, H, W, C, 0, g, ws + l.maps[0], ws + l.plane_acc, n_planes);
    ssim_index_final_kernel<<<(unsigned)cdiv(B, 128), 128, 0, st>>>(ws + l.plane_acc, 1.0f / ((float)Ho * (float)Wo), B, C,
                                                                    out);
    return check_launch("ssim_index");
}

}  // extern "C"
