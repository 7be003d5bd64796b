// Third-generation weight-gradient kernel: tcgen05 kind::f16 with MN-MAJOR operands -- no transposer.
//
// dW[kh][kw][ca][cb] = sum over pixels of P[pixel + (kh, kw)][ca] * Q[pixel][cb]: the pixels are the GEMM K dimension.
// kind::tf32 only takes K-major operands, so conv_tc_wgrad2 transposes every NHWC tile in shared memory (and writes the
// P tile once per kernel tap): round-2 stamps showed it bound by that transposer and by ~10 KB of operand reads per
// MMA (profiles/r02z_wg2_stamps.log), not by the tensor pipe.  kind::f16 takes MN-major tiles: rows = K = pixels,
// 128-byte rows of 64 channels, SWIZZLE_128B -- which is what an NHWC tile IS (probe: scratch/umma_mn16.cu,
// profiles/r02_umma_mn16.log: LBO = distance of the 64-channel panels, SBO = 1024, and a start address shifted by
// whole rows works).  So here:
//   * chunk = 32 consecutive pixels of one image row; its P halo box (KH rows x (32 + KW - 1) pixels) is converted ONCE
//     into fp16 hi / lo panels, and kernel tap (kh, kw) is the same box read through a descriptor shifted by
//     kh * PW + kw rows (the trick of conv_tc_halo.cu, on the K side);
//   * A = [P_hi | P_lo] (two 64-channel panels, M = 128), B = Q_hi, then Q_lo (N = Cb): accumulator rows [0, 64) collect
//     P_hi (Q_hi + Q_lo), rows [64, 128) P_lo (Q_hi + Q_lo); the epilogue adds the two row groups (the lo * lo term it
//     picks up is 2^-22 of the product).  One accumulator of N columns per tap: taps * N <= 512 TMEM columns;
//   * fp16's exponent range: every CTA first reads the pixel range it owns (plus the halo rows) for max |P| and max |Q|
//     and scales each operand by the power of two that brings the maximum into [2^13, 2^14); its partial dW is multiplied
//     by 1 / (s_P s_Q) before the global reduction (exact).  Same 22 significant bits per operand as 3xTF32.
// Per 32-pixel chunk of a 3x3 layer: 36 MMAs of N = Cb (wgrad2: 16 MMAs of N = 216 behind a transposer), ~9 KB of
// producer writes instead of ~140 KB.
#include <stdlib.h>

#include <atomic>

#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace dl4ds {

using namespace tc;

namespace {

constexpr int kW3Threads = 288;         // warps 0-7 producers, then epilogue; warp 8 MMA issuer
constexpr int kW3MaxStages = 6;
constexpr int kW3Chunk = 32;            // pixels per chunk (two K = 16 steps)

struct Wg3Params {
    const float* P; const float* Q; float* dw;
    int p_ld, q_ld;
    int H, W, Ca, Cb, KH, KW, pad_t, pad_l;
    int PW, box_rows;                   // halo box: pixels per row (32 + KW - 1), rows of the box (KH * PW)
    int panel;                          // bytes of one 64-channel panel of the A stage (box rows rounded up, * 128)
    int stage_bytes, stages;            // [P_hi panel | P_lo panel | Q_hi 4 KB | Q_lo 4 KB]
    int Npad, taps, tmem_cols;
    int tiles_x, tiles_per_img, ntiles, tiles_per_cta;
    long long* dbg;                     // optional clock64 stamps of CTA 0 (dl4ds_debug_set_buffer)
};

#define W3_STAMP(it, id)                                                                   \
    do {                                                                                   \
        if (p.dbg != nullptr && blockIdx.x == 0 && (it) < 64 && lane == 0)                 \
            p.dbg[(it) * 16 + (id)] = clock64();                                           \
    } while (0)

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float amax8(const float4& a, const float4& b, float m) {
    m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
    return fmaxf(m, fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
}

// 8 fp32 values * s -> 16 bytes of fp16 hi and 16 bytes of fp16 lo
__device__ __forceinline__ void split8(const float4& a, const float4& b, float s, uint4& hi, uint4& lo) {
    const float v[8] = {a.x * s, a.y * s, a.z * s, a.w * s, b.x * s, b.y * s, b.z * s, b.w * s};
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const __half2 h2 = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
        const float2 hf = __half22float2(h2);
        const __half2 l2 = __floats2half2_rn(v[2 * q] - hf.x, v[2 * q + 1] - hf.y);
        hw[q] = *reinterpret_cast<const uint32_t*>(&h2);
        lw[q] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    hi = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    lo = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

__global__ void __launch_bounds__(kW3Threads, 1) conv_tc_wgrad3_kernel(const Wg3Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[kW3MaxStages];
    __shared__ __align__(8) uint64_t bar_empty[kW3MaxStages];
    __shared__ __align__(8) uint64_t bar_accum;
    __shared__ uint32_t tmem_base_smem;
    __shared__ int2 row_tab[192];       // per box row: {float offset from the box origin, ky << 16 | px}
    __shared__ float red_s[2][8];
    __shared__ float scale_s[2];

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const sm = smem_raw + (smem_base - smem_u32(smem_raw));
    pdl_launch_dependents();
    if (warp == 0) W3_STAMP(0, 12);

    const int t_begin = blockIdx.x * p.tiles_per_cta;
    const int my_tiles = min(p.ntiles, t_begin + p.tiles_per_cta) - t_begin;
    if (my_tiles <= 0) return;

    // ---- prologue: shared memory / TMEM only (programmatic dependent launch: see common.cuh)
    for (int i = threadIdx.x; i < p.box_rows; i += blockDim.x) {
        const int ky = i / p.PW, px = i - ky * p.PW;
        row_tab[i] = make_int2((ky * p.W + px) * p.p_ld, (ky << 16) | px);
    }
    {   // channel padding (and the rows the boxes never reach) stay zero for the whole kernel
        uint4* z = reinterpret_cast<uint4*>(sm);
        const int n16 = p.stages * p.stage_bytes / 16;
        for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(&bar_full[s]), 8);
            mbar_init(smem_u32(&bar_empty[s]), 1);
        }
        mbar_init(smem_u32(&bar_accum), 1);
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(smem_u32(&tmem_base_smem), (uint32_t)p.tmem_cols);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_smem;
    pdl_wait();
    const float* const Pg = pdl_after_wait(p.P);
    const float* const Qg = pdl_after_wait(p.Q);
    if (warp == 0) W3_STAMP(0, 13);

    const int upa = p.Ca >> 3, upb = p.Cb >> 3;        // 8-channel units per pixel
    if (warp < 8) {
        // ===================== producers (256 threads): scales, then chunk after chunk =====================
        const int tid = threadIdx.x;
        // ---- pass 1: max |P| over the chunks this CTA reads (its range widened by the halo rows and one chunk on
        // either side for the halo columns; a neighbouring image only makes the bound more conservative) and max |Q|.
        // The chunks tile the (N, H, W) pixels in order, so both ranges are contiguous runs of pixels; units (8 channels
        // of a pixel) are walked with increments instead of divisions, eight 32-byte loads in flight per thread.
        float mp = 0.0f, mq = 0.0f;
        {
            const int ext = p.tiles_x * max(p.pad_t, p.KH - 1 - p.pad_t) + 1;
            const int c0 = max(0, t_begin - ext), c1 = min(p.ntiles, t_begin + my_tiles + ext);
#pragma unroll
            for (int which = 0; which < 2; ++which) {
                const float* base = which == 0 ? Pg + (int64_t)c0 * kW3Chunk * p.p_ld : Qg + (int64_t)t_begin * kW3Chunk * p.q_ld;
                const int ld = which == 0 ? p.p_ld : p.q_ld, upx = which == 0 ? upa : upb;
                const int npx = (which == 0 ? (c1 - c0) : my_tiles) * kW3Chunk;
                const int total = npx * upx;
                const int dpx = 256 / upx, du = 256 - dpx * upx;        // one step of 256 units
                int px = tid / upx, u = tid - px * upx;
                float m = 0.0f;
                for (int i0 = tid; i0 < total; i0 += 256 * 8) {
                    float4 a[8][2];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        a[j][0] = a[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (i0 + j * 256 < total) {
                            const float4* src = reinterpret_cast<const float4*>(base + (int64_t)px * ld + u * 8);
                            a[j][0] = __ldg(src);
                            a[j][1] = __ldg(src + 1);
                        }
                        px += dpx; u += du;
                        if (u >= upx) { u -= upx; ++px; }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) m = amax8(a[j][0], a[j][1], m);
                }
                if (which == 0) mp = m; else mq = m;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mp = fmaxf(mp, __shfl_xor_sync(0xffffffffu, mp, o));
                mq = fmaxf(mq, __shfl_xor_sync(0xffffffffu, mq, o));
            }
            if (lane == 0) { red_s[0][warp] = mp; red_s[1][warp] = mq; }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mp = red_s[0][0]; mq = red_s[1][0];
#pragma unroll
            for (int w = 1; w < 8; ++w) { mp = fmaxf(mp, red_s[0][w]); mq = fmaxf(mq, red_s[1][w]); }
        }
        const float sp = pow2_scale_for(mp), sq = pow2_scale_for(mq);
        if (tid == 0) { scale_s[0] = sp; scale_s[1] = sq; }
        if (warp == 0) W3_STAMP(0, 14);

        // ---- pass 2: the chunks.  What a thread does is the same for every chunk -- which box row / channel unit, where
        // it lands in the stage, which image offset it reads: worked out once, so the loop body is loads, the hi / lo
        // split and stores.
        const int units_p = p.box_rows * upa, units_q = kW3Chunk * upb;
        constexpr int kUnits = 8;                   // units per thread and chunk (host: (units_p + units_q) <= 256 * 8)
        int u_dst[kUnits], u_src[kUnits], u_pos[kUnits];
#pragma unroll
        for (int j = 0; j < kUnits; ++j) {
            const int i = tid + j * 256;
            u_dst[j] = -1; u_src[j] = 0; u_pos[j] = 0;
            if (i < units_p) {
                const int r = i / upa, u = i - r * upa;
                const int2 e = row_tab[r];
                u_dst[j] = r * 128 + ((u ^ (r & 7)) << 4);
                u_src[j] = e.x + u * 8;
                u_pos[j] = e.y;                     // ky << 16 | px
            } else if (i < units_p + units_q) {
                const int k = i - units_p;
                const int r = k / upb, u = k - r * upb;
                u_dst[j] = 2 * p.panel + r * 128 + ((u ^ (r & 7)) << 4);
                u_src[j] = r * p.q_ld + u * 8;
                u_pos[j] = -1;                      // Q unit: always inside the image
            }
        }
        for (int it = 0; it < my_tiles; ++it) {
            const int s = it % p.stages;
            mbar_wait(smem_u32(&bar_empty[s]), (uint32_t)(((it / p.stages) & 1) ^ 1));
            if (warp == 0) W3_STAMP(it, 0);
            const int tile = t_begin + it;
            const int img = tile / p.tiles_per_img;
            const int trem = tile - img * p.tiles_per_img;
            const int y = trem / p.tiles_x, x0 = (trem - y * p.tiles_x) * kW3Chunk;
            const int by = y - p.pad_t, bx = x0 - p.pad_l;                 // box origin (may lie outside the image)
            uint8_t* const stage = sm + (size_t)s * p.stage_bytes;
            const float* pbase = Pg + (((int64_t)img * p.H + by) * p.W + bx) * p.p_ld;
            const float* qbase = Qg + (((int64_t)img * p.H + y) * p.W + x0) * p.q_ld;
            float4 a[kUnits][2];
#pragma unroll
            for (int j = 0; j < kUnits; ++j) {
                a[j][0] = a[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                const bool isq = u_pos[j] < 0;
                const int ky = u_pos[j] >> 16, px = u_pos[j] & 0xffff;
                const bool inside = isq || ((unsigned)(by + ky) < (unsigned)p.H && (unsigned)(bx + px) < (unsigned)p.W);
                if (u_dst[j] >= 0 && inside) {
                    const float4* src = reinterpret_cast<const float4*>((isq ? qbase : pbase) + u_src[j]);
                    a[j][0] = __ldg(src);
                    a[j][1] = __ldg(src + 1);
                }
            }
#pragma unroll
            for (int j = 0; j < kUnits; ++j) {
                uint4 hi, lo;
                const bool isq = u_pos[j] < 0;
                split8(a[j][0], a[j][1], isq ? sq : sp, hi, lo);
                if (u_dst[j] >= 0) {
                    *reinterpret_cast<uint4*>(stage + u_dst[j]) = hi;
                    *reinterpret_cast<uint4*>(stage + u_dst[j] + (isq ? 4096 : p.panel)) = lo;
                }
            }
            fence_proxy_async_smem();
            if (warp == 0) W3_STAMP(it, 1);
            mbar_arrive_warp(smem_u32(&bar_full[s]));
        }

        // ===================== epilogue (the same eight warps) =====================
        mbar_wait(smem_u32(&bar_accum), 0);
        tc_fence_after();
        if (warp == 0) W3_STAMP(1, 12);
        // TMEM lane quadrant = warp % 4: quadrants 2, 3 hold the P_lo rows of channels [0, 64); they park their
        // values in shared memory (every stage is free now), quadrants 0, 1 add them and reduce into dw
        const int q = warp & 3, half = warp >> 2;           // the two warps of a quadrant take alternate 16-column blocks
        float* const stg = reinterpret_cast<float*>(sm);    // [tap][Npad columns][64 rows]: lane = row, so a warp's accesses are 32 consecutive words
        const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16);
        const int ncol = p.taps * p.Npad;
        if (q >= 2) {
            const int m = (q - 2) * 32 + lane;
            for (int c0 = half * 16; c0 < ncol; c0 += 32) {
                float v[16];
                tmem_ld16(taddr + (uint32_t)c0, v);
                const int tap = c0 / p.Npad, n0 = c0 - tap * p.Npad;
#pragma unroll
                for (int j = 0; j < 16; ++j) stg[((size_t)tap * p.Npad + n0 + j) * 64 + m] = v[j];
            }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (q < 2) {
            const int m = q * 32 + lane;
            const float unscale = (1.0f / scale_s[0]) * (1.0f / scale_s[1]);
            for (int c0 = half * 16; c0 < ncol; c0 += 32) {
                float v[16];
                tmem_ld16(taddr + (uint32_t)c0, v);
                const int tap = c0 / p.Npad, n0 = c0 - tap * p.Npad;
                if (m < p.Ca) {
                    float* dst = p.dw + ((int64_t)tap * p.Ca + m) * p.Cb + n0;
                    const float* lo = stg + ((size_t)tap * p.Npad + n0) * 64 + m;
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        if (n0 + j < p.Cb)
                            red_add_v4(dst + j, (v[j] + lo[j * 64]) * unscale, (v[j + 1] + lo[(j + 1) * 64]) * unscale,
                                       (v[j + 2] + lo[(j + 2) * 64]) * unscale, (v[j + 3] + lo[(j + 3) * 64]) * unscale);
                    }
                }
            }
        }
        if (warp == 0) W3_STAMP(1, 13);
        tc_fence_before();
    } else {
        // ===================== MMA issuer (warp 8; one elected lane issues) =====================
        uint32_t elected;
        asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(elected));
        const bool leader = elected != 0;
        const uint32_t idesc = make_idesc_f16(128, p.Npad) | (1u << 15) | (1u << 16);      // A and B MN-major
        const uint64_t tmpl_a = make_smem_desc(0, (uint32_t)p.panel, 1024, kLayoutSw128);
        const uint64_t tmpl_b = make_smem_desc(0, 4096, 1024, kLayoutSw128);
        const uint64_t a_jump = (uint64_t)(((p.PW - (p.KW - 1)) * 128) >> 4);              // last tap of a kernel row -> first of the next
        for (int it = 0; it < my_tiles; ++it) {
            const int s = it % p.stages;
            mbar_wait(smem_u32(&bar_full[s]), (uint32_t)((it / p.stages) & 1));
            tc_fence_after();
            W3_STAMP(it, 4);
            const uint32_t a_addr = smem_base + (uint32_t)(s * p.stage_bytes);
            const uint32_t b_addr = a_addr + 2u * (uint32_t)p.panel;
            const uint64_t da0 = tmpl_a + (uint64_t)((a_addr & 0x3FFFFu) >> 4);
            const uint64_t db_hi = tmpl_b + (uint64_t)((b_addr & 0x3FFFFu) >> 4);
            const uint64_t db_lo = db_hi + (4096u >> 4);
            const uint32_t acc0 = it > 0 ? 1u : 0u;
            // the whole warp walks the taps (uniform registers, increments only); the elected lane issues -- a single
            // thread's scalar chain would pace the MMAs (conv_tc_halo.cu)
            uint64_t da = da0;
            uint32_t td = tmem_d;
            int kw = 0;
            for (int tap = 0; tap < p.taps; ++tap) {
                if (leader) {
                    umma_f16(td, da, db_hi, idesc, acc0);                                   // pixels [0, 16)
                    umma_f16(td, da, db_lo, idesc, 1u);
                    umma_f16(td, da + (2048u >> 4), db_hi + (2048u >> 4), idesc, 1u);       // pixels [16, 32)
                    umma_f16(td, da + (2048u >> 4), db_lo + (2048u >> 4), idesc, 1u);
                }
                td += (uint32_t)p.Npad;
                if (++kw == p.KW) { kw = 0; da += a_jump; } else { da += 8u; }             // next tap: one row (128 B) on
            }
            if (leader) umma_commit(smem_u32(&bar_empty[s]));
            W3_STAMP(it, 5);
            __syncwarp();
        }
        if (leader) umma_commit(smem_u32(&bar_accum));
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_d, (uint32_t)p.tmem_cols);
}

}  // namespace

extern std::atomic<long long> g_tc_launches;
static long long* g_w3_dbg = nullptr;
void wgrad3_set_debug_buffer(long long* p) { g_w3_dbg = p; }

// DL4DS_E_UNSUPPORTED outside the domain (then conv_tc_wgrad2 / the other kernels of api.cu run)
int conv2d_wgrad_tc3(const WgradArgs& a, int math_mode, cudaStream_t st) {
    static const bool disabled = [] { const char* e = getenv("DL4DS_WGRAD3"); return e && e[0] == '0'; }();
    if (disabled) return DL4DS_E_UNSUPPORTED;
    if (math_mode != DL4DS_MATH_TF32X3) return DL4DS_E_UNSUPPORTED;        // the fp32-parity mode (F16X3 maps onto it in api.cu)
    if (dl4ds_device_is_sm100() != 1) return DL4DS_E_UNSUPPORTED;
    if (a.stride != 1 || a.Hp != a.Hq || a.Wp != a.Wq) return DL4DS_E_UNSUPPORTED;
    if (a.Wq % kW3Chunk) return DL4DS_E_UNSUPPORTED;
    if (a.Ca % 8 || a.Cb % 8 || a.Ca > 64 || a.Cb > 64 || a.p_ld % 4 || a.q_ld % 4) return DL4DS_E_UNSUPPORTED;
    // Narrow layers stay on conv_tc_wgrad2: this kernel issues taps * 4 MMAs per chunk whatever the widths (the ~44-cycle
    // floor of a small-N MMA), which only pays once the stacked-rows kernel's transposer is the larger cost.  Measured
    // (CTA 0, clk, 64 x 32 x 32): 16 -> 16 43.6k against 33.4k, 24 -> 24 52k / 40k, 32 -> 32 53k / 57k, 40 -> 40 56k / 63k,
    // 48 -> 48 65k / 71k; 48 -> 32 at 64 x 64: 164k / 237k.  (DL4DS_WGRAD3_MIN_C overrides the threshold.)
    static const int min_c = [] { const char* e = getenv("DL4DS_WGRAD3_MIN_C"); return e ? atoi(e) : 32; }();
    if (a.Ca < min_c || a.Cb < min_c) return DL4DS_E_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(a.P) & 15) || (reinterpret_cast<uintptr_t>(a.Q) & 15) ||
        (reinterpret_cast<uintptr_t>(a.dw) & 15))
        return DL4DS_E_UNSUPPORTED;
    const int taps = a.KH * a.KW;
    if (taps == 1 || a.KH > 5 || a.KW > 5) return DL4DS_E_UNSUPPORTED;
    // (Wider layers could be split into channel-group roles -- <= 64 input channels, taps x N <= 512 accumulator columns
    //  per role; tried in round 2: every role converts the P box again and issues its own taps x 4 MMAs per chunk, the
    //  48 -> 192 sub-pixel layer took 100 us against 92 us on conv_tc_wgrad2.  Not kept.)
    Wg3Params p;
    p.Npad = (a.Cb + 15) & ~15;
    if (taps * p.Npad > 512) return DL4DS_E_UNSUPPORTED;
    p.P = a.P; p.Q = a.Q; p.dw = a.dw; p.p_ld = a.p_ld; p.q_ld = a.q_ld;
    p.H = a.Hq; p.W = a.Wq; p.Ca = a.Ca; p.Cb = a.Cb; p.KH = a.KH; p.KW = a.KW; p.pad_t = a.pad_t; p.pad_l = a.pad_l;
    p.PW = kW3Chunk + a.KW - 1;
    p.box_rows = a.KH * p.PW;
    if (p.box_rows > 192 || p.box_rows * (a.Ca / 8) + kW3Chunk * (a.Cb / 8) > 256 * 8) return DL4DS_E_UNSUPPORTED;
    p.panel = ((p.box_rows + 7) & ~7) * 128;
    p.panel = (p.panel + 1023) & ~1023;
    p.stage_bytes = 2 * p.panel + 8192;
    p.taps = taps;
    p.dbg = g_w3_dbg;
    int cols = 32;
    while (cols < taps * p.Npad) cols *= 2;
    p.tmem_cols = cols;
    p.tiles_x = a.Wq / kW3Chunk;
    p.tiles_per_img = p.tiles_x * a.Hq;
    p.ntiles = a.N * p.tiles_per_img;
    const int budget = 216 * 1024;
    int stages = budget / p.stage_bytes;
    if (stages > kW3MaxStages) stages = kW3MaxStages;
    const size_t epi = (size_t)taps * 64 * p.Npad * 4;          // epilogue staging reuses the stages
    if (stages < 2 || (size_t)stages * p.stage_bytes < epi) return DL4DS_E_UNSUPPORTED;
    p.stages = stages;
    int grid = p.ntiles < kNumSMs ? p.ntiles : kNumSMs;
    p.tiles_per_cta = (p.ntiles + grid - 1) / grid;
    grid = (p.ntiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
    const size_t smem = (size_t)stages * p.stage_bytes + 1024;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(conv_tc_wgrad3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024));
        attr_done = true;
    }
    launch_pdl(2, conv_tc_wgrad3_kernel, dim3(grid), dim3(kW3Threads), smem, st, p);
    g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    return check_launch("conv_tc_wgrad3_kernel");
}

}  // namespace dl4ds
