// CUDA-core fp32 implicit-GEMM convolution family (DL4DS_MATH_FP32).
//
// Exact-fp32 path for every Conv2D / Conv2DTranspose forward, input-gradient and weight-gradient
// the DL4DS graphs need (reference call sites: dl4ds/models/blocks.py:49-61,208,299,414-416,479,
// 508-516; sp_postups.py:134,156).  The tcgen05 kernels in conv_tc.cu take over the tensor-bound
// shapes; this file stays the path for narrow layers (Cin or Cout < 8, strided / fractionally
// strided taps) and is the fp32 parity anchor for the tensor-core kernels.
//
// GEMM view: M = N*Ho*Wo output pixels, N = Cout, K = KH*KW*Cin, K walked tap by tap in chunks of
// 8 input channels.  256 threads per CTA, register tile TM x TN per thread, A (pixels x 8 ch) and
// B (8 ch x BN) staged through shared memory with a register prefetch of the next chunk.
#include "common.cuh"

namespace dl4ds {

// `splits` > 1 (split-K): blockIdx.z takes a contiguous range of the (tap, channel-chunk) loop and ADDS its raw partial
// sums into y (zeroed by the launcher); bias / residual / activation are applied afterwards by conv_finish_kernel.
// Used when M x Cout yields only a handful of CTAs but K is long (the 4x4 / 8x8 levels of the U-Net, sp_preups.py:
// 230-315: 256 pixels x 256 channels x K=2304 ran on 4-16 CTAs for 340 us).
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256) conv_fwd_kernel(const ConvArgs p, int splits) {
    constexpr int BK = 8;
    constexpr int APT = BM * BK / 256;   // A floats per thread per chunk (4 or 8)
    constexpr int TPP = BK / APT;        // threads per pixel (2 or 1)
    constexpr int NTX = BN / TN;
    constexpr int BPT = (BK * BN + 255) / 256;
    static_assert((BM / TM) * NTX == 256, "tile/thread mismatch");
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];

    const int tid = threadIdx.x;
    const int tx = tid % NTX, ty = tid / NTX;
    const int n0 = blockIdx.y * BN;

    // the pixel this thread stages into As
    const int la_m = tid / TPP, la_k = (tid % TPP) * APT;
    const int m_glob = blockIdx.x * BM + la_m;
    const bool m_ok = m_glob < p.M;
    int pn = 0, base_y = 0, base_x = 0;
    if (m_ok) {
        pn = m_glob / p.HoWo;
        const int r = m_glob - pn * p.HoWo;
        const int oy = r / p.Wo, ox = r - oy * p.Wo;
        base_y = oy * p.stride - p.pad_t;
        base_x = ox * p.stride - p.pad_l;
    }
    const int cpt = (p.Cin + BK - 1) / BK;
    const int nchunks_all = p.KH * p.KW * cpt;
    const int ntaps = p.KH * p.KW;
    const int per = (nchunks_all + splits - 1) / splits;
    const int chunk0 = blockIdx.z * per;
    const int nchunks = min(per, nchunks_all - chunk0);
    if (nchunks <= 0) return;

    float ra[APT];
    float rb[BPT];

    auto load_chunk = [&](int chunk) {
        const int tap = chunk / cpt;
        const int cc = (chunk - tap * cpt) * BK;
        const int kh = tap / p.KW, kw = tap - kh * p.KW;
        // ---- A
        int iy = base_y + kh, ix = base_x + kw;
        bool ok = m_ok && iy >= 0 && ix >= 0;
        if (p.up > 1) {
            ok = ok && (iy % p.up == 0) && (ix % p.up == 0);
            iy /= p.up; ix /= p.up;
        }
        ok = ok && iy < p.H && ix < p.W;
        const int c0 = cc + la_k;
#pragma unroll
        for (int j = 0; j < APT; ++j) ra[j] = 0.0f;
        if (ok) {
            const float* src = p.x + ((int64_t)(pn * p.H + iy) * p.W + ix) * p.x_ld + c0;
            if (p.vec) {
#pragma unroll
                for (int j = 0; j < APT; j += 4) {
                    if (c0 + j < p.Cin) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(src + j));
                        ra[j] = v.x; ra[j + 1] = v.y; ra[j + 2] = v.z; ra[j + 3] = v.w;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < APT; ++j)
                    if (c0 + j < p.Cin) ra[j] = __ldg(src + j);
            }
        }
        // ---- B
#pragma unroll
        for (int j = 0; j < BPT; ++j) {
            const int i = tid + j * 256;
            float v = 0.0f;
            if (i < BK * BN) {
                const int k = i / BN, n = i - k * BN;
                const int c = cc + k, co = n0 + n;
                if (c < p.Cin && co < p.Cout) {
                    if (p.wmode == DL4DS_W_HWIO)
                        v = __ldg(p.w + ((int64_t)tap * p.Cin + c) * p.Cout + co);
                    else
                        v = __ldg(p.w + ((int64_t)(ntaps - 1 - tap) * p.Cout + co) * p.Cin + c);
                }
            }
            rb[j] = v;
        }
    };
    auto store_chunk = [&]() {
#pragma unroll
        for (int j = 0; j < APT; ++j) As[la_k + j][la_m] = ra[j];
#pragma unroll
        for (int j = 0; j < BPT; ++j) {
            const int i = tid + j * 256;
            if (i < BK * BN) {
                const int k = i / BN, n = i - k * BN;
                Bs[k][n] = rb[j];
            }
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    load_chunk(chunk0);
    store_chunk();
    __syncthreads();
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        if (chunk + 1 < nchunks) load_chunk(chunk0 + chunk + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(&As[k][ty * TM + i]);
                a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
        if (chunk + 1 < nchunks) {
            store_chunk();
            __syncthreads();
        }
    }

    // ---- epilogue: bias, residual, activation, (depth_to_space) store
    const int r = p.d2s_r;
    const int Cd = (r > 1) ? p.Cout / (r * r) : p.Cout;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = blockIdx.x * BM + ty * TM + i;
        if (m >= p.M) continue;
        int n = 0, oy = 0, ox = 0;
        if (r > 1) {
            n = m / p.HoWo;
            const int rem = m - n * p.HoWo;
            oy = rem / p.Wo; ox = rem - oy * p.Wo;
        }
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int co = n0 + tx * TN + j;
            if (co >= p.Cout) continue;
            float v = acc[i][j];
            if (splits > 1) {
                atomicAdd(p.y + (int64_t)m * p.y_ld + co, v);
                continue;
            }
            if (p.bias) v += __ldg(p.bias + co);
            if (p.res) v += __ldg(p.res + (int64_t)m * p.res_ld + co);
            v = apply_act(v, p.act);
            if (r == 1) {
                float* dst = p.y + (int64_t)m * p.y_ld + co;
                *dst = p.beta ? (*dst + v) : v;
            } else {
                const int g = co / Cd, c = co - g * Cd;
                const int di = g / r, dj = g - di * r;
                const int64_t pix = ((int64_t)(n * p.Ho * r + oy * r + di)) * (p.Wo * r) + ox * r + dj;
                p.y[pix * p.y_ld + c] = v;
            }
        }
    }
}

// y = act(y + bias + res) over (M, Cout) with pitch: second pass of the split-K path
__global__ void conv_finish_kernel(float* __restrict__ y, int y_ld, const float* __restrict__ bias,
                                   const float* __restrict__ res, int res_ld, int64_t M, int Cout, int act) {
    const int64_t total = M * Cout;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / Cout;
        const int co = (int)(i - m * Cout);
        float v = y[m * y_ld + co];
        if (bias) v += __ldg(bias + co);
        if (res) v += __ldg(res + m * res_ld + co);
        y[m * y_ld + co] = apply_act(v, act);
    }
}

template <int BM, int BN, int TM, int TN>
static int launch_conv(const ConvArgs& a, cudaStream_t st) {
    dim3 grid((unsigned)cdiv(a.M, BM), (unsigned)cdiv(a.Cout, BN));
    const int nchunks = a.KH * a.KW * ((a.Cin + 7) / 8);
    const int ctas = (int)(grid.x * grid.y);
    int splits = 1;
    if (ctas * 4 <= kNumSMs && nchunks >= 32 && a.d2s_r <= 1 && !a.beta) {
        splits = (2 * kNumSMs) / ctas;
        if (splits > nchunks / 8) splits = nchunks / 8;      // >= 8 chunks (64 K rows) per split
    }
    if (splits > 1) {
        grid.z = (unsigned)splits;
        if (cudaMemset2DAsync(a.y, (size_t)a.y_ld * 4, 0, (size_t)a.Cout * 4, (size_t)a.M, st) != cudaSuccess)
            return check_launch("conv_fwd split-K memset");
        conv_fwd_kernel<BM, BN, TM, TN><<<grid, 256, 0, st>>>(a, splits);
        int rc = check_launch("conv_fwd_kernel");
        if (rc) return rc;
        const int64_t total = (int64_t)a.M * a.Cout;
        const int blocks = (int)((total + 255) / 256 > 8 * kNumSMs ? 8 * kNumSMs : (total + 255) / 256);
        conv_finish_kernel<<<blocks, 256, 0, st>>>(a.y, a.y_ld, a.bias, a.res, a.res_ld, a.M, a.Cout, a.act);
        return check_launch("conv_finish_kernel");
    }
    conv_fwd_kernel<BM, BN, TM, TN><<<grid, 256, 0, st>>>(a, 1);
    return check_launch("conv_fwd_kernel");
}

int conv2d_fwd_simt(const ConvArgs& a, cudaStream_t st) {
    const int c = a.Cout;
    if (c <= 4) return launch_conv<256, 4, 4, 1>(a, st);
    if (c <= 8) return launch_conv<256, 8, 4, 2>(a, st);
    if (c <= 16) return launch_conv<128, 16, 4, 2>(a, st);
    if (c <= 24) return launch_conv<128, 24, 4, 3>(a, st);
    if (c <= 32) return launch_conv<128, 32, 4, 4>(a, st);
    if (c <= 40) return launch_conv<128, 40, 4, 5>(a, st);
    if (c <= 48) return launch_conv<128, 48, 8, 3>(a, st);
    return launch_conv<128, 64, 8, 4>(a, st);
}

// -------------------------------------------------------------------------------------------------
// weight gradient: dw[(tap,a)][b] += sum_q P[shift_tap(q)][a] * Q[q][b]
// GEMM view: M = KH*KW*Ca, N = Cb, K = N*Hq*Wq pixels (split across gridDim.z, fp32 atomics).
// -------------------------------------------------------------------------------------------------
template <int TMW, int TBW, int RA, int RB>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const WgradArgs p) {
    constexpr int BP = 32;
    constexpr int NTB = TBW / RB;
    static_assert((TMW / RA) * NTB == 256, "tile/thread mismatch");
    __shared__ __align__(16) float Ps[BP][TMW + 4];
    __shared__ __align__(16) float Qs[BP][TBW + 4];
    __shared__ short m_kh[TMW], m_kw[TMW];
    __shared__ int m_a[TMW];
    __shared__ int s_n[BP], s_y[BP], s_x[BP];

    const int tid = threadIdx.x;
    const int tb = tid % NTB, ta = tid / NTB;
    const int m0 = blockIdx.x * TMW, b0 = blockIdx.y * TBW;
    for (int i = tid; i < TMW; i += 256) {
        const int m = m0 + i;
        if (m < p.Mw) {
            const int tap = m / p.Ca;
            m_a[i] = m - tap * p.Ca;
            m_kh[i] = (short)(tap / p.KW);
            m_kw[i] = (short)(tap % p.KW);
        } else {
            m_a[i] = -1; m_kh[i] = 0; m_kw[i] = 0;
        }
    }
    float acc[RA][RB];
#pragma unroll
    for (int i = 0; i < RA; ++i)
#pragma unroll
        for (int j = 0; j < RB; ++j) acc[i][j] = 0.0f;

    const int64_t chunk0 = (int64_t)blockIdx.z * p.chunks_per_split;
    const int HqWq = p.Hq * p.Wq;
    for (int ch = 0; ch < p.chunks_per_split; ++ch) {
        const int64_t qbase = (chunk0 + ch) * BP;
        if (qbase >= p.NQ) break;
        __syncthreads();   // previous chunk's tiles fully consumed (also covers the m_* table init)
        if (tid < BP) {
            const int64_t q = qbase + tid;
            if (q < p.NQ) {
                const int n = (int)(q / HqWq);
                const int rem = (int)(q - (int64_t)n * HqWq);
                const int oy = rem / p.Wq, ox = rem - oy * p.Wq;
                s_n[tid] = n;
                s_y[tid] = oy * p.stride - p.pad_t;
                s_x[tid] = ox * p.stride - p.pad_l;
            } else {
                s_n[tid] = -1; s_y[tid] = 0; s_x[tid] = 0;
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int i = tid; i < BP * TMW; i += 256) {
            const int pp = i / TMW, mm = i - pp * TMW;
            float v = 0.0f;
            const int a = m_a[mm], n = s_n[pp];
            if (a >= 0 && n >= 0) {
                const int iy = s_y[pp] + m_kh[mm], ix = s_x[pp] + m_kw[mm];
                if (iy >= 0 && iy < p.Hp && ix >= 0 && ix < p.Wp)
                    v = __ldg(p.P + ((int64_t)(n * p.Hp + iy) * p.Wp + ix) * p.p_ld + a);
            }
            Ps[pp][mm] = v;
        }
#pragma unroll 4
        for (int i = tid; i < BP * TBW; i += 256) {
            const int pp = i / TBW, bb = i - pp * TBW;
            float v = 0.0f;
            const int b = b0 + bb;
            if (s_n[pp] >= 0 && b < p.Cb) v = __ldg(p.Q + (qbase + pp) * p.q_ld + b);
            Qs[pp][bb] = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int pp = 0; pp < BP; ++pp) {
            float a[RA], b[RB];
#pragma unroll
            for (int i = 0; i < RA; ++i) a[i] = Ps[pp][ta * RA + i];
#pragma unroll
            for (int j = 0; j < RB; ++j) b[j] = Qs[pp][tb * RB + j];
#pragma unroll
            for (int i = 0; i < RA; ++i)
#pragma unroll
                for (int j = 0; j < RB; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < RA; ++i) {
        const int m = m0 + ta * RA + i;
        if (m >= p.Mw) continue;
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const int b = b0 + tb * RB + j;
            if (b < p.Cb) atomicAdd(p.dw + (int64_t)m * p.Cb + b, acc[i][j]);
        }
    }
}

template <int TMW, int TBW, int RA, int RB>
static int launch_wgrad(WgradArgs a, cudaStream_t st) {
    const int64_t chunks = cdiv(a.NQ, 32);
    const int64_t tiles = cdiv(a.Mw, TMW) * cdiv(a.Cb, TBW);
    int64_t nsplit = (4 * kNumSMs + tiles - 1) / tiles;
    if (nsplit > chunks) nsplit = chunks;
    if (nsplit < 1) nsplit = 1;
    a.chunks_per_split = (int)cdiv(chunks, nsplit);
    nsplit = cdiv(chunks, a.chunks_per_split);
    dim3 grid((unsigned)cdiv(a.Mw, TMW), (unsigned)cdiv(a.Cb, TBW), (unsigned)nsplit);
    conv_wgrad_kernel<TMW, TBW, RA, RB><<<grid, 256, 0, st>>>(a);
    return check_launch("conv_wgrad_kernel");
}

int conv2d_wgrad_simt(const WgradArgs& a, cudaStream_t st) {
    if (a.Cb <= 4) return launch_wgrad<256, 4, 4, 1>(a, st);
    if (a.Cb <= 8) return launch_wgrad<256, 8, 4, 2>(a, st);
    if (a.Cb <= 16) return launch_wgrad<128, 16, 4, 2>(a, st);
    if (a.Cb <= 32) return launch_wgrad<64, 32, 4, 2>(a, st);
    return launch_wgrad<64, 64, 4, 4>(a, st);
}

}  // namespace dl4ds
