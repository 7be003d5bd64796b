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
#include <cuda_fp16.h>
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

// -------------------------------------------------------------------------------------------------
// The 3-term mode on fp16 operands (round 2).  Measured on B200 (profiles/r02_mma_sync_rate.log): one
// mma.sync.m16n8k8.tf32 and one mma.sync.m16n8k16.f16 both issue every 8 clocks per SM sub-partition, so an fp16
// MMA carries twice the MACs: K = 16 holds the 8 input channels of TWO taps and the nine taps take 5 MMAs instead
// of 9.  The operands are split ONCE, when the tile is staged (x * s = hi + lo, both fp16, s a power of two that
// puts the tile's maximum in [2^13, 2^14); the weights likewise, per tensor), so the inner loop is fragment loads
// and MMAs only: acc = hi*hi + (lo*hi + hi*lo), 22 mantissa bits like 3xTF32, un-scaled exactly in the epilogue.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ float thin_pow2_scale(float amax) {      // 2^(13 - floor(log2 amax)); 1 for 0 / inf / nan
    const int e = (int)((__float_as_uint(amax) >> 23) & 0xffu);
    if (e == 0 || e == 255) return 1.0f;
    int se = 127 + 13 - (e - 127);
    se = se < 1 ? 1 : (se > 254 ? 254 : se);
    return __uint_as_float((uint32_t)se << 23);
}

__device__ __forceinline__ void split_h2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - f.x, x1 - f.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ void mma_f16_16x8x16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ld_nc_f2(const float* p, float2& v) {
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
}
__device__ __forceinline__ void ld_f2(const float* p, float2& v) {
    asm volatile("ld.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
}

__device__ __forceinline__ float block_amax_256(float m, float* red8) {     // red8: 8 floats of shared memory
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red8[threadIdx.x >> 5] = m;
    __syncthreads();
    float r = red8[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) r = fmaxf(r, red8[i]);
    return r;
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) thin_conv_f16_kernel(ConvArgs p, int TW, int tiles_x, int tiles_y) {
    pdl_launch_dependents();    // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    p.x = pdl_after_wait(p.x);
    p.res = pdl_after_wait(p.res);
    p.mask_y = pdl_after_wait(p.mask_y);
    constexpr int TH = 8;
    extern __shared__ __align__(16) uint32_t smw[];          // [hi | lo] planes, (TH+2) rows x pitch words, 4 words = one pixel
    __shared__ float red8[8];
    __shared__ float dbs[8];
    // dgrad launches with the producer's epilogue-backward fused (dl4ds_conv2d_dgrad_fused): dz = (acc [+ old]) *
    // act'(y_producer), dbias += column sums of dz; no bias / residual / activation of its own
    const bool fusedbw = p.mask_y != nullptr || p.dbias != nullptr;
    if (threadIdx.x < 8) dbs[threadIdx.x] = 0.f;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int pitch = (TW + 2) * 4;
    const int plane = (TH + 2) * pitch;
    // B fragments of the 5 tap pairs: b0 = W[tap 2pr][ci 2t, 2t+1][co g], b1 = W[tap 2pr+1][ci 2t, 2t+1][co g]
    float wv[5][4];
    float wmax = 0.f;
#pragma unroll
    for (int pr = 0; pr < 5; ++pr)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int tap = 2 * pr + (j >> 1), ci = 2 * t + (j & 1);
            float w = 0.f;
            if (tap < 9)
                w = (p.wmode == DL4DS_W_HWIO) ? __ldg(p.w + (tap * 8 + ci) * 8 + g) : __ldg(p.w + ((8 - tap) * 8 + g) * 8 + ci);
            wv[pr][j] = w;
            wmax = fmaxf(wmax, fabsf(w));
        }
#pragma unroll
    for (int o = 16; o; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    const float sw = thin_pow2_scale(wmax);
    uint32_t bh[5][2], bl[5][2];
#pragma unroll
    for (int pr = 0; pr < 5; ++pr)
#pragma unroll
        for (int h = 0; h < 2; ++h) split_h2(wv[pr][2 * h] * sw, wv[pr][2 * h + 1] * sw, bh[pr][h], bl[pr][h]);
    const int tile = blockIdx.x;
    const int img = tile / (tiles_x * tiles_y);
    const int trem = tile - img * tiles_x * tiles_y;
    const int ty = trem / tiles_x, tx = trem - ty * tiles_x;
    const int y0 = ty * TH, x0 = tx * TW;
    float sx;
    {   // input tile with halo: all of a thread's (<= 11) 16-byte loads are in flight together; the tile maximum gives
        // the scale, then every value is split once and stored as 4 + 4 halfs
        const int cols = TW + 2, total = (TH + 2) * cols * 2;
        float4 v[kMaxLoads];
        float m = 0.f;
        // (row, pixel) of item i = tid + 256 k advance by 128 pixels per step: one division per thread, not two per load
        // (ncu r02fin_thin16: 275 instructions per 16-pixel tile and warp, a third of them this index arithmetic)
        const int c4 = tid & 1;
        const int r0 = (tid >> 1) / cols, px0 = (tid >> 1) - r0 * cols;
        const float* xin = p.x + (int64_t)img * p.H * p.W * p.x_ld + c4 * 4;
        {
            int r = r0, px = px0;
#pragma unroll
            for (int k = 0; k < kMaxLoads; ++k) {
                const int gy = y0 + r - p.pad_t, gx = x0 + px - p.pad_l;
                v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < TH + 2 && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)
                    v[k] = __ldg(reinterpret_cast<const float4*>(xin + (int64_t)(gy * p.W + gx) * p.x_ld));
                m = fmaxf(fmaxf(m, fmaxf(fabsf(v[k].x), fabsf(v[k].y))), fmaxf(fabsf(v[k].z), fabsf(v[k].w)));
                px += 128;
                while (px >= cols) { px -= cols; ++r; }
            }
        }
        sx = thin_pow2_scale(block_amax_256(m, red8));
        {
            int r = r0, px = px0;
#pragma unroll
            for (int k = 0; k < kMaxLoads; ++k) {
                if (r < TH + 2) {
                    uint2 hi, lo;
                    split_h2(v[k].x * sx, v[k].y * sx, hi.x, lo.x);
                    split_h2(v[k].z * sx, v[k].w * sx, hi.y, lo.y);
                    uint32_t* dst = smw + r * pitch + px * 4 + c4 * 2;
                    *reinterpret_cast<uint2*>(dst) = hi;
                    *reinterpret_cast<uint2*>(dst + plane) = lo;
                }
                px += 128;
                while (px >= cols) { px -= cols; ++r; }
            }
        }
        (void)total;
    }
    __syncthreads();
    const int oy = y0 + warp;
    if (oy >= p.H && !fusedbw) return;
    const float inv = (1.0f / sx) * (1.0f / sw);
    const float bias0 = p.bias ? __ldg(p.bias + 2 * t) : 0.f, bias1 = p.bias ? __ldg(p.bias + 2 * t + 1) : 0.f;
    float bs0 = 0.f, bs1 = 0.f;
    // row bases of the epilogue's tensors (64-bit once; the loop adds 32-bit pixel offsets)
    const int64_t rowpix = ((int64_t)img * p.H + oy) * p.W + x0 + g;
    const float* const e_src = fusedbw ? p.mask_y : p.res;
    const int e_ld = fusedbw ? p.mask_ld : p.res_ld;
    const float* const e_row = e_src ? e_src + rowpix * e_ld + 2 * t : nullptr;
    float* const y_row = p.y + rowpix * p.y_ld + 2 * t;
    for (int m0 = 0; m0 < TW && oy < p.H; m0 += 16) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f}, accx[4] = {0.f, 0.f, 0.f, 0.f};
        // the epilogue's global operands (residual / producer output / old value) are requested BEFORE the MMAs of this
        // pixel tile (asm volatile keeps them there): loaded at the point of use they cost a memory round trip per tile
        // and per warp -- r02cfg5f: 11 us without, 17 us with a residual, 26 us with mask + old at 4 x 256 x 256
        float2 e_a[2], e_b[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            e_a[h] = e_b[h] = make_float2(0.f, 0.f);
            if (e_row) ld_nc_f2(e_row + (m0 + 8 * h) * e_ld, e_a[h]);
            if (p.beta) ld_f2(y_row + (m0 + 8 * h) * p.y_ld, e_b[h]);
        }
#pragma unroll
        for (int pr = 0; pr < 5; ++pr) {
            const int ta = 2 * pr, tb = 2 * pr + 1;
            const uint32_t* ra = smw + (warp + ta / 3) * pitch + (m0 + g + ta % 3) * 4 + t;
            uint32_t ah[4], al[4];
            ah[0] = ra[0]; ah[1] = ra[32];
            al[0] = ra[plane]; al[1] = ra[plane + 32];
            if (tb < 9) {
                const uint32_t* rb = smw + (warp + tb / 3) * pitch + (m0 + g + tb % 3) * 4 + t;
                ah[2] = rb[0]; ah[3] = rb[32];
                al[2] = rb[plane]; al[3] = rb[plane + 32];
            } else {
                ah[2] = ah[3] = al[2] = al[3] = 0u;
            }
            mma_f16_16x8x16(accx, al, bh[pr][0], bh[pr][1]);
            mma_f16_16x8x16(accx, ah, bl[pr][0], bl[pr][1]);
            mma_f16_16x8x16(acc, ah, bh[pr][0], bh[pr][1]);
        }
        // C fragment: acc[0..1] = pixel m0+g, channels 2t, 2t+1; acc[2..3] = pixel m0+g+8
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float2 o = make_float2((acc[2 * h] + accx[2 * h]) * inv + bias0, (acc[2 * h + 1] + accx[2 * h + 1]) * inv + bias1);
            float2* const dstf = reinterpret_cast<float2*>(y_row + (m0 + 8 * h) * p.y_ld);
            if (fusedbw) {
                o.x += e_b[h].x; o.y += e_b[h].y;
                if (p.mask_y) {
                    o.x *= act_grad_from_out(e_a[h].x, p.mask_act); o.y *= act_grad_from_out(e_a[h].y, p.mask_act);
                }
                *dstf = o;
                bs0 += o.x; bs1 += o.y;
                continue;
            }
            o.x += e_a[h].x; o.y += e_a[h].y;
            o.x = apply_act(o.x, p.act); o.y = apply_act(o.y, p.act);
            o.x += e_b[h].x; o.y += e_b[h].y;
            *dstf = o;
        }
    }
    if (p.dbias != nullptr) {           // (uniform per launch; every warp of the block arrives here in this mode)
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            bs0 += __shfl_xor_sync(0xffffffffu, bs0, o);
            bs1 += __shfl_xor_sync(0xffffffffu, bs1, o);
        }
        if (lane < 4) { atomicAdd(&dbs[2 * lane], bs0); atomicAdd(&dbs[2 * lane + 1], bs1); }
        __syncthreads();
        // two 16-byte reductions per block (the arena keeps every tensor 16-byte aligned): at batch 4 every block of the
        // launch lands on the same 32 bytes within a few microseconds, and same-address atomics serialise in L2
        if ((reinterpret_cast<uintptr_t>(p.dbias) & 15) == 0) {
            if (tid < 2)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.dbias + 4 * tid), "f"(dbs[4 * tid]),
                             "f"(dbs[4 * tid + 1]), "f"(dbs[4 * tid + 2]), "f"(dbs[4 * tid + 3]) : "memory");
        } else if (tid < 8) {
            atomicAdd(p.dbias + tid, dbs[tid]);
        }
    }
}

static bool thin_f16_enabled() {
    static const bool on = [] { const char* e = getenv("DL4DS_THIN_F16"); return !(e && e[0] == '0'); }();
    return on;
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
    if ((a.mask_y || a.dbias) && !(math_mode == DL4DS_MATH_TF32X3 && thin_f16_enabled())) return DL4DS_E_UNSUPPORTED;
    if (a.mask_y && (a.mask_ld % 2 || (reinterpret_cast<uintptr_t>(a.mask_y) & 7))) return DL4DS_E_UNSUPPORTED;
    const int TW = a.W > 128 ? 128 : a.W;
    const int tiles_x = a.W / TW, tiles_y = (a.H + 7) / 8;
    const size_t smem = (size_t)10 * ((TW + 2) * 8 + 8) * 4;
    if (math_mode == DL4DS_MATH_TF32X3 && thin_f16_enabled()) {
        // 2 blocks per SM (118 registers, no spills) measured faster than 3 (85 registers, 60 B spilled): cfg5 6.15 vs 6.56 ms
        static const bool minb2 = [] { const char* e = getenv("DL4DS_THIN_F16_MINB"); return !(e && e[0] == '3'); }();
        const size_t smem16 = (size_t)2 * 10 * (TW + 2) * 4 * 4;
        if (minb2)
            launch_pdl(4, thin_conv_f16_kernel<2>, dim3(a.N * tiles_x * tiles_y), dim3(256), smem16, st, a, TW, tiles_x, tiles_y);
        else
            launch_pdl(4, thin_conv_f16_kernel<3>, dim3(a.N * tiles_x * tiles_y), dim3(256), smem16, st, a, TW, tiles_x, tiles_y);
    }
    else if (math_mode == DL4DS_MATH_TF32X3)
        launch_pdl(4, thin_conv_mma_kernel<true>, dim3(a.N * tiles_x * tiles_y), dim3(256), smem, st, a, TW, tiles_x, tiles_y);
    else
        launch_pdl(4, thin_conv_mma_kernel<false>, dim3(a.N * tiles_x * tiles_y), dim3(256), smem, st, a, TW, tiles_x, tiles_y);
    return check_launch("thin_conv_mma_kernel");
}

// the shapes whose dgrad can carry the producer's epilogue-backward (the fp16 3-term kernel only)
bool conv2d_thin_fused_dgrad_supported(int N, int H, int W, int KH, int KW, int math_mode) {
    static const bool disabled = [] { const char* e = getenv("DL4DS_THIN_NO_MMA"); return e && e[0] == '1'; }();
    if (math_mode == DL4DS_MATH_F16X3) math_mode = DL4DS_MATH_TF32X3;
    static const bool off = [] { const char* e = getenv("DL4DS_THIN_FUSED"); return e && e[0] == '0'; }();
    if (off || disabled || math_mode != DL4DS_MATH_TF32X3 || !thin_f16_enabled()) return false;
    if (KH != 3 || KW != 3 || W % 32 || (W > 128 && W % 128)) return false;
    return (int64_t)N * H * W >= 16384;
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

// -------------------------------------------------------------------------------------------------
// The weight gradient on fp16 operands: K = 16 consecutive pixels of a row per MMA (half as many MMAs as the tf32
// kernel above), M = 16 = the 8 input channels of two taps, N = the 8 output channels.  The tiles are staged
// channel-major (a line = one channel of one tile row, pixel pairs packed in 32-bit words) and split once into
// hi / lo halves under per-tile power-of-two scales; taps with an odd horizontal offset take their pixel pairs from
// two neighbouring words (PRMT).  Scales differ from tile to tile, so the MMA accumulators are folded into fp32
// sums (un-scaled, exact) after every tile.
// -------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int line_pitch_words(int px) {       // >= ceil(px / 2), congruent 4 (mod 32): conflict-free fragments
    int n = (px + 1) / 2;
    return ((n - 4 + 31) / 32) * 32 + 4;
}

// one channel-major staging step: the float4 `v` (4 channels c4*4.. of pixel px) of this lane and of lane ^ 2 (pixel
// px ^ 1) become 2 channels x (px & ~1, px | 1) packed pairs, split and stored.  Every lane must call it (shuffles).
__device__ __forceinline__ void stage_pairs(const float4& v, float s, bool odd, bool store, uint32_t* hi_line0, int plane,
                                            int pitch) {
    const float s0 = odd ? v.x : v.z, s1 = odd ? v.y : v.w;
    const float r0 = __shfl_xor_sync(0xffffffffu, s0, 2), r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
    if (!store) return;
    uint32_t h0, l0, h1, l1;
    if (!odd) {         // channels +0, +1: (own, partner)
        split_h2(v.x * s, r0 * s, h0, l0);
        split_h2(v.y * s, r1 * s, h1, l1);
    } else {            // channels +2, +3: (partner, own)
        split_h2(r0 * s, v.z * s, h0, l0);
        split_h2(r1 * s, v.w * s, h1, l1);
    }
    hi_line0[0] = h0; hi_line0[plane] = l0;
    hi_line0[pitch] = h1; hi_line0[pitch + plane] = l1;
}

__global__ void __launch_bounds__(256, 2) thin_wgrad_f16_kernel(const float* __restrict__ P, int p_ld, const float* __restrict__ Q,
                                                             int q_ld, float* __restrict__ dw, int N, int H, int W, int pad_t,
                                                             int pad_l, int TW, int tiles_x, int tiles_y) {
    pdl_launch_dependents();    // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    P = pdl_after_wait(P);
    Q = pdl_after_wait(Q);
    constexpr int TH = 8;
    extern __shared__ __align__(16) uint32_t smw[];
    __shared__ float red[576];
    __shared__ float redp[8], redq[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int LP = line_pitch_words(TW + 2), LQ = line_pitch_words(TW);
    const int planeP = (TH + 2) * 8 * LP, planeQ = TH * 8 * LQ;
    uint32_t* Ph = smw;                                      // [hi | lo] x (TH+2) rows x 8 channels x LP words
    uint32_t* Qh = smw + 2 * planeP;                         // [hi | lo] x TH rows x 8 channels x LQ words
    float sum[5][4];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) sum[i][j] = 0.f;
    for (int i = tid; i < 576; i += 256) red[i] = 0.f;
    const int ntiles = N * tiles_x * tiles_y;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int img = tile / (tiles_x * tiles_y);
        const int trem = tile - img * tiles_x * tiles_y;
        const int ty = trem / tiles_x, tx = trem - ty * tiles_x;
        const int y0 = ty * TH, x0 = tx * TW;
        __syncthreads();                                     // previous tile fully consumed
        float sp, sq;
        {
            const int cols = TW + 2, total = (TH + 2) * cols * 2;
            float4 v[kMaxLoads];
            float m = 0.f;
            // (row, pixel) of item tid + 256 k advance by 128 pixels per step: one division per thread and tile
            const int c4 = tid & 1;
            const int r0 = (tid >> 1) / cols, px0 = (tid >> 1) - r0 * cols;
            const float* Pin = P + (int64_t)img * H * W * p_ld + c4 * 4;
            {
                int r = r0, px = px0;
#pragma unroll
                for (int k = 0; k < kMaxLoads; ++k) {
                    const int gy = y0 + r - pad_t, gx = x0 + px - pad_l;
                    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < TH + 2 && gy >= 0 && gy < H && gx >= 0 && gx < W)
                        v[k] = __ldg(reinterpret_cast<const float4*>(Pin + (int64_t)(gy * W + gx) * p_ld));
                    m = fmaxf(fmaxf(m, fmaxf(fabsf(v[k].x), fabsf(v[k].y))), fmaxf(fabsf(v[k].z), fabsf(v[k].w)));
                    px += 128;
                    while (px >= cols) { px -= cols; ++r; }
                }
            }
            sp = thin_pow2_scale(block_amax_256(m, redp));
            {
                int r = r0, px = px0;
#pragma unroll
                for (int k = 0; k < kMaxLoads; ++k) {
                    const bool odd = px & 1;
                    stage_pairs(v[k], sp, odd, r < TH + 2, Ph + (r * 8 + c4 * 4 + (odd ? 2 : 0)) * LP + (px >> 1), planeP, LP);
                    px += 128;
                    while (px >= cols) { px -= cols; ++r; }
                }
            }
            (void)total;
        }
        {
            const int totq = TH * TW * 2;
            float4 v[8];
            float m = 0.f;
            const int c4 = tid & 1;
            const int r0 = (tid >> 1) / TW, px0 = (tid >> 1) - r0 * TW;
            const float* Qin = Q + ((int64_t)(img * H + y0) * W + x0) * q_ld + c4 * 4;
            {
                int r = r0, px = px0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < TH && y0 + r < H) v[k] = __ldg(reinterpret_cast<const float4*>(Qin + (int64_t)(r * W + px) * q_ld));
                    m = fmaxf(fmaxf(m, fmaxf(fabsf(v[k].x), fabsf(v[k].y))), fmaxf(fabsf(v[k].z), fabsf(v[k].w)));
                    px += 128;
                    while (px >= TW) { px -= TW; ++r; }
                }
            }
            sq = thin_pow2_scale(block_amax_256(m, redq));
            {
                int r = r0, px = px0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const bool odd = px & 1;
                    stage_pairs(v[k], sq, odd, r < TH, Qh + (r * 8 + c4 * 4 + (odd ? 2 : 0)) * LQ + (px >> 1), planeQ, LQ);
                    px += 128;
                    while (px >= TW) { px -= TW; ++r; }
                }
            }
            (void)totq;
        }
        __syncthreads();
        float acc[5][4];
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        const uint32_t* qline = Qh + (warp * 8 + g) * LQ + t;
        for (int k0 = 0; k0 < TW; k0 += 16) {
            // B fragment: b0 = Q[px k0+2t, k0+2t+1][co g], b1 = the same 8 pixels further
            const uint32_t bh0 = qline[k0 >> 1], bh1 = qline[(k0 >> 1) + 4];
            const uint32_t bl0 = qline[planeQ + (k0 >> 1)], bl1 = qline[planeQ + (k0 >> 1) + 4];
#pragma unroll
            for (int pr = 0; pr < 5; ++pr) {
                // A fragment: rows 0-7 = input channels of tap 2*pr (a0, a2), rows 8-15 = of tap 2*pr+1 (a1, a3; tap 9: zeros)
                uint32_t ah[4], al[4];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int tap = 2 * pr + h;
                    if (tap < 9) {
                        const int kh = tap / 3, kw = tap % 3;
                        const uint32_t* pl = Ph + ((warp + kh) * 8 + g) * LP + (k0 >> 1) + t;
                        if (kw == 1) {
                            ah[h] = __byte_perm(pl[0], pl[1], 0x5432);
                            ah[h + 2] = __byte_perm(pl[4], pl[5], 0x5432);
                            al[h] = __byte_perm(pl[planeP], pl[planeP + 1], 0x5432);
                            al[h + 2] = __byte_perm(pl[planeP + 4], pl[planeP + 5], 0x5432);
                        } else {
                            const int o = kw >> 1;
                            ah[h] = pl[o];
                            ah[h + 2] = pl[o + 4];
                            al[h] = pl[planeP + o];
                            al[h + 2] = pl[planeP + o + 4];
                        }
                    } else {
                        ah[h] = ah[h + 2] = al[h] = al[h + 2] = 0u;
                    }
                }
                mma_f16_16x8x16(acc[pr], al, bh0, bh1);
                mma_f16_16x8x16(acc[pr], ah, bl0, bl1);
                mma_f16_16x8x16(acc[pr], ah, bh0, bh1);
            }
        }
        const float inv = (1.0f / sp) * (1.0f / sq);
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sum[i][j] = fmaf(acc[i][j], inv, sum[i][j]);
    }
    // C fragment of pair pr: [0..1] = dw[tap 2pr][ci g][co 2t, 2t+1], [2..3] = dw[tap 2pr+1][ci g][co 2t, 2t+1]
    __syncthreads();
#pragma unroll
    for (int pr = 0; pr < 5; ++pr)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int tap = 2 * pr + h;
            if (tap < 9) {
                atomicAdd(&red[(tap * 8 + g) * 8 + 2 * t], sum[pr][2 * h]);
                atomicAdd(&red[(tap * 8 + g) * 8 + 2 * t + 1], sum[pr][2 * h + 1]);
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
    const bool f16 = math_mode == DL4DS_MATH_TF32X3 && thin_f16_enabled();
    const size_t smem = f16 ? (size_t)2 * (10 * 8 * line_pitch_words(TW + 2) + 8 * 8 * line_pitch_words(TW)) * 4
                            : (size_t)(10 * ((TW + 2) * 8 + 8) + 8 * (TW * 8 + 8)) * 4;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(thin_wgrad_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
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
    if (f16)
        launch_pdl(4, thin_wgrad_f16_kernel, dim3(grid), dim3(256), smem, st, w.P, w.p_ld, w.Q, w.q_ld, w.dw, w.N, w.Hq, w.Wq,
                   w.pad_t, w.pad_l, TW, tiles_x, tiles_y);
    else if (math_mode == DL4DS_MATH_TF32X3)
        launch_pdl(4, thin_wgrad_mma_kernel<true>, dim3(grid), dim3(256), smem, st, w.P, w.p_ld, w.Q, w.q_ld, w.dw, w.N, w.Hq,
                   w.Wq, w.pad_t, w.pad_l, TW, tiles_x, tiles_y);
    else
        launch_pdl(4, thin_wgrad_mma_kernel<false>, dim3(grid), dim3(256), smem, st, w.P, w.p_ld, w.Q, w.q_ld, w.dw, w.N, w.Hq,
                   w.Wq, w.pad_t, w.pad_l, TW, tiles_x, tiles_y);
    return check_launch("thin_wgrad_mma_kernel");
}

}  // namespace dl4ds
