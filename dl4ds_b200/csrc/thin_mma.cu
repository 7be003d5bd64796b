// Warp-level tensor-core kernels for the 8-channel HR tail (ConvBlock_tail / ConvBlock_out inputs at
// 128 x 128, sp_postups.py:205-212): 3x3 convolution 8 -> 8 forward / input gradient and its weight gradient.
//
// These layers are too narrow for tcgen05 (N = 8, K = 8 per tap: every MMA would re-read its operands from
// shared memory for 512 MACs per pixel) and on CUDA cores they are issue-bound (ncu r01d: 63 % issue-active with
// the FMA pipe at 38 %, 55 us for 604 MFMA).  One `mma.sync.m16n8k8.tf32` is exactly one kernel tap of 16 pixels
// (M = 16 pixels, N = 8 output channels, K = 8 input channels): the operands live in REGISTERS, so the 3xTF32
// split costs three ALU ops per element and no shared-memory round trip, and the weights of all nine taps stay in
// 36 registers per lane for the whole kernel.  3xTF32: acc += lo*hi + hi*lo + hi*hi (hi = round-to-nearest tf32).
// The exact-fp32 math mode keeps the CUDA-core kernels of thin.cu.
#include <stdlib.h>

#include "common.cuh"

namespace dl4ds {

__device__ __forceinline__ float tf32_rn(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const float (&a)[4], float b0, float b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
          "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

constexpr int kMaxLoads = 11;      // ceil(10 rows x 130 pixels x 2 float4 / 256 threads)

// shared-memory position (in floats) of channel c of halo pixel px: the two 4-channel halves of a pixel swap
// places on every other group of 4 pixels, which makes the fragment loads (8 pixels x 4 channels) conflict-free
__device__ __forceinline__ int px_word(int px, int c) { return px * 8 + ((((c >> 2) ^ (px >> 2)) & 1) << 2) + (c & 3); }

// -------------------------------------------------------------------------------------------------
// y = act(conv3x3(x, w) + bias + res) [+= y], 8 -> 8 channels, stride 1, 'same'.  Block = 8 warps = 8 rows of a
// TW-pixel-wide tile; warp = one row, 16 pixels per MMA tile.
// -------------------------------------------------------------------------------------------------
template <bool X3>
__global__ void __launch_bounds__(256, 3) thin_conv_mma_kernel(ConvArgs p, int TW, int tiles_x, int tiles_y) {
    pdl_launch_dependents();    // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    p.x = pdl_after_wait(p.x);
    p.res = pdl_after_wait(p.res);
    constexpr int TH = 8;
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int pitch = (TW + 2) * 8 + 8;
    // B fragments of the 9 taps: b0 = W[tap][ci = t][co = g], b1 = W[tap][ci = t + 4][co = g]
    float bh[9][2], bl[9][2];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ci = t + 4 * h;
            const float w = (p.wmode == DL4DS_W_HWIO) ? __ldg(p.w + (tap * 8 + ci) * 8 + g)
                                                       : __ldg(p.w + ((8 - tap) * 8 + g) * 8 + ci);
            bh[tap][h] = X3 ? tf32_rn(w) : w;
            bl[tap][h] = w - bh[tap][h];
        }
    const int tile = blockIdx.x;
    const int img = tile / (tiles_x * tiles_y);
    const int trem = tile - img * tiles_x * tiles_y;
    const int ty = trem / tiles_x, tx = trem - ty * tiles_x;
    const int y0 = ty * TH, x0 = tx * TW;
    {   // input tile with halo: rows y0-pad_t .. +TH+2, cols x0-pad_l .. +TW+2
        // all of a thread's (<= 11) 16-byte loads are issued before the first store: the tile load is the
        // latency-critical phase of this kernel (one DRAM round trip instead of ten)
        const int cols = TW + 2, total = (TH + 2) * cols * 2;
        float4 v[kMaxLoads];
#pragma unroll
        for (int k = 0; k < kMaxLoads; ++k) {
            const int i = tid + k * 256;
            const int c4 = i & 1, px = (i >> 1) % cols, r = (i >> 1) / cols;
            const int gy = y0 + r - p.pad_t, gx = x0 + px - p.pad_l;
            v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < total && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)
                v[k] = __ldg(reinterpret_cast<const float4*>(p.x + ((int64_t)(img * p.H + gy) * p.W + gx) * p.x_ld) + c4);
        }
#pragma unroll
        for (int k = 0; k < kMaxLoads; ++k) {
            const int i = tid + k * 256;
            const int c4 = i & 1, px = (i >> 1) % cols, r = (i >> 1) / cols;
            if (i < total) *reinterpret_cast<float4*>(sm + (size_t)r * pitch + px_word(px, c4 * 4)) = v[k];
        }
    }
    __syncthreads();
    const int oy = y0 + warp;
    if (oy >= p.H) return;
    const float bias0 = p.bias ? __ldg(p.bias + 2 * t) : 0.f, bias1 = p.bias ? __ldg(p.bias + 2 * t + 1) : 0.f;
    for (int m0 = 0; m0 < TW; m0 += 16) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const float* row = sm + (size_t)(warp + kh) * pitch;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int pa = m0 + g + kw, pb = pa + 8;
                float a[4];
                a[0] = row[px_word(pa, t)];
                a[1] = row[px_word(pb, t)];
                a[2] = row[px_word(pa, t + 4)];
                a[3] = row[px_word(pb, t + 4)];
                const int tap = kh * 3 + kw;
                if (X3) {
                    float ah[4], al[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { ah[i] = tf32_rn(a[i]); al[i] = a[i] - ah[i]; }
                    mma_tf32_16x8x8(acc, al, bh[tap][0], bh[tap][1]);
                    mma_tf32_16x8x8(acc, ah, bl[tap][0], bl[tap][1]);
                    mma_tf32_16x8x8(acc, ah, bh[tap][0], bh[tap][1]);
                } else {
                    mma_tf32_16x8x8(acc, a, bh[tap][0], bh[tap][1]);
                }
            }
        }
        // C fragment: acc[0..1] = pixel m0+g, channels 2t, 2t+1; acc[2..3] = pixel m0+g+8
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t pix = ((int64_t)img * p.H + oy) * p.W + x0 + m0 + g + 8 * h;
            float2 o = make_float2(acc[2 * h] + bias0, acc[2 * h + 1] + bias1);
            if (p.res) {
                const float2 r = __ldg(reinterpret_cast<const float2*>(p.res + pix * p.res_ld + 2 * t));
                o.x += r.x; o.y += r.y;
            }
            o.x = apply_act(o.x, p.act); o.y = apply_act(o.y, p.act);
            float2* dst = reinterpret_cast<float2*>(p.y + pix * p.y_ld + 2 * t);
            if (p.beta) { const float2 old = *dst; o.x += old.x; o.y += old.y; }
            *dst = o;
        }
    }
}

// DL4DS_E_UNSUPPORTED outside the domain (then thin.cu's CUDA-core kernel runs)
int conv2d_fwd_thin_mma(const ConvArgs& a, int math_mode, cudaStream_t st) {
    static const bool disabled = [] { const char* e = getenv("DL4DS_THIN_NO_MMA"); return e && e[0] == '1'; }();
    if (math_mode == DL4DS_MATH_F16X3) math_mode = DL4DS_MATH_TF32X3;
    if (disabled || (math_mode != DL4DS_MATH_TF32X3 && math_mode != DL4DS_MATH_TF32)) return DL4DS_E_UNSUPPORTED;
    if (a.KH != 3 || a.KW != 3 || a.stride != 1 || a.up != 1 || a.Ho != a.H || a.Wo != a.W || a.d2s_r > 1)
        return DL4DS_E_UNSUPPORTED;
    if (a.Cin != 8 || a.Cout != 8 || !a.vec) return DL4DS_E_UNSUPPORTED;
    if (a.W % 32 || (a.W > 128 && a.W % 128)) return DL4DS_E_UNSUPPORTED;
    if (a.y_ld % 2 || (reinterpret_cast<uintptr_t>(a.y) & 7)) return DL4DS_E_UNSUPPORTED;
    if (a.res && (a.res_ld % 2 || (reinterpret_cast<uintptr_t>(a.res) & 7))) return DL4DS_E_UNSUPPORTED;
    if ((int64_t)a.N * a.H * a.W < 16384) return DL4DS_E_UNSUPPORTED;
    const int TW = a.W > 128 ? 128 : a.W;
    const int tiles_x = a.W / TW, tiles_y = (a.H + 7) / 8;
    const size_t smem = (size_t)10 * ((TW + 2) * 8 + 8) * 4;
    if (math_mode == DL4DS_MATH_TF32X3)
        launch_pdl(4, thin_conv_mma_kernel<true>, dim3(a.N * tiles_x * tiles_y), dim3(256), smem, st, a, TW, tiles_x, tiles_y);
    else
        launch_pdl(4, thin_conv_mma_kernel<false>, dim3(a.N * tiles_x * tiles_y), dim3(256), smem, st, a, TW, tiles_x, tiles_y);
    return check_launch("thin_conv_mma_kernel");
}

// -------------------------------------------------------------------------------------------------
// dw[kh][kw][ci][co] += sum_px P[px + (kh,kw) - pad][ci] * Q[px][co], 8 x 8 channels, 3x3.  The reduction runs
// over pixels: K = 8 consecutive pixels of a row per MMA; M = 16 carries the 8 input channels of TWO taps
// (rows 0-7: tap 2i, rows 8-15: tap 2i+1), N = the 8 output channels -> 5 MMAs (x3: 15) per 8 pixels, and every
// lane ends up owning 18 distinct elements of dw (no intra-warp reduction).  Block = 8 warps = 8 rows of a tile;
// blocks are persistent and merge once into dw (shared-memory atomics, then one global atomic per element).
// -------------------------------------------------------------------------------------------------
template <bool X3>
__global__ void __launch_bounds__(256, 2) thin_wgrad_mma_kernel(const float* __restrict__ P, int p_ld, const float* __restrict__ Q,
                                                             int q_ld, float* __restrict__ dw, int N, int H, int W, int pad_t,
                                                             int pad_l, int TW, int tiles_x, int tiles_y) {
    pdl_launch_dependents();    // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    P = pdl_after_wait(P);
    Q = pdl_after_wait(Q);
    constexpr int TH = 8;
    extern __shared__ __align__(16) float sm[];
    __shared__ float red[576];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int ppitch = (TW + 2) * 8 + 8, qpitch = TW * 8 + 8;
    float* Ps = sm;                                          // (TH + 2) x ppitch
    float* Qs = sm + (TH + 2) * ppitch;                      // TH x qpitch
    float acc[5][4];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int i = tid; i < 576; i += 256) red[i] = 0.f;
    const int ntiles = N * tiles_x * tiles_y;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int img = tile / (tiles_x * tiles_y);
        const int trem = tile - img * tiles_x * tiles_y;
        const int ty = trem / tiles_x, tx = trem - ty * tiles_x;
        const int y0 = ty * TH, x0 = tx * TW;
        __syncthreads();                                     // previous tile fully consumed
        {
            const int cols = TW + 2, total = (TH + 2) * cols * 2;
            const int totq = TH * TW * 2;
            float4 v[kMaxLoads];
#pragma unroll
            for (int k = 0; k < kMaxLoads; ++k) {
                const int i = tid + k * 256;
                const int c4 = i & 1, px = (i >> 1) % cols, r = (i >> 1) / cols;
                const int gy = y0 + r - pad_t, gx = x0 + px - pad_l;
                v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < total && gy >= 0 && gy < H && gx >= 0 && gx < W)
                    v[k] = __ldg(reinterpret_cast<const float4*>(P + ((int64_t)(img * H + gy) * W + gx) * p_ld) + c4);
            }
#pragma unroll
            for (int k = 0; k < kMaxLoads; ++k) {
                const int i = tid + k * 256;
                const int c4 = i & 1, px = (i >> 1) % cols, r = (i >> 1) / cols;
                if (i < total) *reinterpret_cast<float4*>(Ps + (size_t)r * ppitch + px * 8 + c4 * 4) = v[k];
            }
            float4 (&vq)[kMaxLoads] = v;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = tid + k * 256;
                const int c4 = i & 1, px = (i >> 1) % TW, r = (i >> 1) / TW;
                vq[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < totq && y0 + r < H)
                    vq[k] = __ldg(reinterpret_cast<const float4*>(Q + ((int64_t)(img * H + y0 + r) * W + x0 + px) * q_ld) + c4);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = tid + k * 256;
                const int c4 = i & 1, px = (i >> 1) % TW, r = (i >> 1) / TW;
                if (i < totq) *reinterpret_cast<float4*>(Qs + (size_t)r * qpitch + px * 8 + c4 * 4) = vq[k];
            }
        }
        __syncthreads();
        const float* qrow = Qs + (size_t)warp * qpitch;
        for (int k0 = 0; k0 < TW; k0 += 8) {
            // B fragment: b0 = Q[px k0+t][co g], b1 = Q[px k0+t+4][co g]
            float b[2], bhi[2], blo[2];
            b[0] = qrow[(k0 + t) * 8 + g];
            b[1] = qrow[(k0 + t + 4) * 8 + g];
#pragma unroll
            for (int i = 0; i < 2; ++i) { bhi[i] = X3 ? tf32_rn(b[i]) : b[i]; blo[i] = b[i] - bhi[i]; }
#pragma unroll
            for (int pr = 0; pr < 5; ++pr) {
                // A fragment: rows 0-7 = input channels of tap 2*pr, rows 8-15 = of tap 2*pr+1 (tap 9 does not exist: zeros)
                const int ta = 2 * pr, tb = 2 * pr + 1;
                const float* ra = Ps + (size_t)(warp + ta / 3) * ppitch + (k0 + ta % 3) * 8 + g;
                float a[4];
                a[0] = ra[t * 8];
                a[2] = ra[(t + 4) * 8];
                if (tb < 9) {
                    const float* rb = Ps + (size_t)(warp + tb / 3) * ppitch + (k0 + tb % 3) * 8 + g;
                    a[1] = rb[t * 8];
                    a[3] = rb[(t + 4) * 8];
                } else {
                    a[1] = 0.f; a[3] = 0.f;
                }
                if (X3) {
                    float ah[4], al[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { ah[i] = tf32_rn(a[i]); al[i] = a[i] - ah[i]; }
                    mma_tf32_16x8x8(acc[pr], al, bhi[0], bhi[1]);
                    mma_tf32_16x8x8(acc[pr], ah, blo[0], blo[1]);
                    mma_tf32_16x8x8(acc[pr], ah, bhi[0], bhi[1]);
                } else {
                    mma_tf32_16x8x8(acc[pr], a, bhi[0], bhi[1]);
                }
            }
        }
    }
    // C fragment of pair pr: acc[0..1] = dw[tap 2pr][ci g][co 2t, 2t+1], acc[2..3] = dw[tap 2pr+1][ci g][co 2t, 2t+1]
    __syncthreads();
#pragma unroll
    for (int pr = 0; pr < 5; ++pr)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int tap = 2 * pr + h;
            if (tap < 9) {
                atomicAdd(&red[(tap * 8 + g) * 8 + 2 * t], acc[pr][2 * h]);
                atomicAdd(&red[(tap * 8 + g) * 8 + 2 * t + 1], acc[pr][2 * h + 1]);
            }
        }
    __syncthreads();
    for (int i = tid; i < 576; i += 256) atomicAdd(dw + i, red[i]);
}

// DL4DS_E_UNSUPPORTED outside the domain (then thin.cu's CUDA-core kernel runs)
int conv2d_wgrad_thin_mma(const WgradArgs& w, int math_mode, cudaStream_t st) {
    static const bool disabled = [] { const char* e = getenv("DL4DS_THIN_NO_MMA"); return e && e[0] == '1'; }();
    if (math_mode == DL4DS_MATH_F16X3) math_mode = DL4DS_MATH_TF32X3;
    if (disabled || (math_mode != DL4DS_MATH_TF32X3 && math_mode != DL4DS_MATH_TF32)) return DL4DS_E_UNSUPPORTED;
    if (w.KH != 3 || w.KW != 3 || w.stride != 1 || w.Hp != w.Hq || w.Wp != w.Wq) return DL4DS_E_UNSUPPORTED;
    if (w.Ca != 8 || w.Cb != 8) return DL4DS_E_UNSUPPORTED;
    if (w.Wq % 32 || (w.Wq > 128 && w.Wq % 128) || w.Hq % 8) return DL4DS_E_UNSUPPORTED;
    if (w.p_ld % 4 || w.q_ld % 4 || (reinterpret_cast<uintptr_t>(w.P) & 15) || (reinterpret_cast<uintptr_t>(w.Q) & 15))
        return DL4DS_E_UNSUPPORTED;
    if (w.NQ < 16384) return DL4DS_E_UNSUPPORTED;
    const int TW = w.Wq > 128 ? 128 : w.Wq;
    const int tiles_x = w.Wq / TW, tiles_y = w.Hq / 8;
    const size_t smem = (size_t)(10 * ((TW + 2) * 8 + 8) + 8 * (TW * 8 + 8)) * 4;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(thin_wgrad_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        cudaFuncSetAttribute(thin_wgrad_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        attr = true;
    }
    int blocks_per_sm = (int)((200 * 1024) / (smem + 4096));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    if (blocks_per_sm > 3) blocks_per_sm = 3;
    int grid = kNumSMs * blocks_per_sm;
    const int ntiles = w.N * tiles_x * tiles_y;
    if (grid > ntiles) grid = ntiles;
    if (math_mode == DL4DS_MATH_TF32X3)
        launch_pdl(4, thin_wgrad_mma_kernel<true>, dim3(grid), dim3(256), smem, st, w.P, w.p_ld, w.Q, w.q_ld, w.dw, w.N, w.Hq,
                   w.Wq, w.pad_t, w.pad_l, TW, tiles_x, tiles_y);
    else
        launch_pdl(4, thin_wgrad_mma_kernel<false>, dim3(grid), dim3(256), smem, st, w.P, w.p_ld, w.Q, w.q_ld, w.dw, w.N, w.Hq,
                   w.Wq, w.pad_t, w.pad_l, TW, tiles_x, tiles_y);
    return check_launch("thin_wgrad_mma_kernel");
}

}  // namespace dl4ds
