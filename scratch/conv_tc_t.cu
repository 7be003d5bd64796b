// tcgen05 implicit-GEMM convolution, "transposed" orientation, 3xTF32 only (forward of narrow layers and every
// dgrad whose output is narrow: blocks.py:49-61,208,299,414-416; sp_postups.py:134,156).
//
// Why a second kernel: one kind::tf32 M=128 K=8 MMA costs ~119 cycles for ANY N <= 128 (171 at N=256,
// scratch/umma_rate.cu), and clock64 stamps of conv_tc_fwd_kernel (scratch/fwd_stamps.py) show its MMA warp
// blocked on exactly that: with pixels on M (128 per tile) and Cout <= 64 on N, a K=8 slice costs
// 2 MMAs x 119 cycles per 128 pixels.  Here the roles are swapped:
//   A (M = 128 rows)  = the weights of one (tap, channel chunk): rows [0,64) the tf32 hi parts of the <= 64 output
//                       channels, rows [64,128) their lo parts -- both MMAs of a K slice use this one operand;
//   B (N = 256 rows)  = 256 output pixels (a BH x BW patch) x kc input channels, K-major exactly as TMA delivers
//                       the tap-shifted NHWC box (out-of-bounds = zero padding); split in place into hi / lo;
//   D (128 x 256 fp32 in TMEM, two buffers = all 512 columns): lanes [0,64) collect W_hi X_hi + W_hi X_lo,
//                       lanes [64,128) W_lo X_hi (+ W_lo X_lo); the epilogue adds the two lane groups through a
//                       shared-memory exchange.
// => 2 MMAs x 171 cycles per 256 pixels and K slice: 0.72x the tensor time of the pixel-major kernel.
//
// Warp roles (576 threads): warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2-9 epilogue (lane quadrant
// = warp % 4: quadrants 0,1 own the hi rows and write the output, quadrants 2,3 publish the lo rows), warps 10-17
// hi/lo splitter of the pixel tile.
#include <stdlib.h>

#include "tc_common.cuh"

namespace dl4ds {

using namespace tc;

constexpr int kTtThreads = 576;
constexpr int kTtMaxStages = 8;
constexpr int kTtExPitch = 260;           // floats per exchange row (256 pixels + 4: conflict-free 16-byte rows)

struct TcTParams {
    const float* wp;                      // packed weight images [tap][chunk][128 rows][kc] (swizzled)
    const float* bias;
    const float* res;
    float* y;
    int res_ld, y_ld;
    int H, W, Cin, Cout;
    int KW, ntaps, pad_t, pad_l;
    int BW, BH, tiles_x, tiles_per_img, ntiles;
    int kc, span, nchunks;
    uint32_t layout;
    int act, d2s_r, beta;
    int stages, stage_bytes, x_bytes, w_bytes;
    int ex_off;                           // byte offset of the lo-row exchange buffer
};

// weights -> [tap][chunk][128][kc]: row n < Cout = tf32 hi of output channel n, row 64 + n = its lo part, rest 0
__global__ void pack_weights_t_kernel(const float* __restrict__ w, float* __restrict__ out, int taps, int Cin, int Cout,
                                      int kc, int nchunks, int wmode) {
    const int upr = kc / 4;
    const int64_t total = (int64_t)taps * nchunks * 128 * upr;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int u = (int)(idx % upr);
        const int row = (int)((idx / upr) % 128);
        const int blk = (int)(idx / ((int64_t)upr * 128));
        const int tap = blk / nchunks, ch = blk - tap * nchunks;
        const int n = row & 63;
        const bool lo = row >= 64;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = ch * kc + u * 4 + j;
            float x = 0.0f;
            if (n < Cout && c < Cin) {
                x = (wmode == DL4DS_W_HWIO) ? __ldg(w + ((int64_t)tap * Cin + c) * Cout + n)
                                             : __ldg(w + ((int64_t)(taps - 1 - tap) * Cout + n) * Cin + c);
                const float h = tf32_rna(x);
                x = lo ? x - h : h;
            }
            v[j] = x;
        }
        const int us = swizzle_unit(u, row, kc * 4);
        *reinterpret_cast<float4*>(out + ((int64_t)blk * 128 + row) * kc + us * 4) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

__global__ void __launch_bounds__(kTtThreads, 1)
conv_tc_fwdT_kernel(const __grid_constant__ CUtensorMap tmap_x, const TcTParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[kTtMaxStages];
    __shared__ __align__(8) uint64_t bar_conv[kTtMaxStages];
    __shared__ __align__(8) uint64_t bar_empty[kTtMaxStages];
    __shared__ __align__(8) uint64_t bar_tfull[2];
    __shared__ __align__(8) uint64_t bar_tempty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float bias_s[64];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    const int nit = p.ntaps * p.nchunks;
    for (int i = threadIdx.x; i < 64; i += blockDim.x) bias_s[i] = (p.bias && i < p.Cout) ? __ldg(p.bias + i) : 0.0f;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(&bar_full[s]), 1);
            mbar_init(smem_u32(&bar_conv[s]), 8);
            mbar_init(smem_u32(&bar_empty[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&bar_tfull[b]), 1);
            mbar_init(smem_u32(&bar_tempty[b]), 8);
        }
        fence_barrier_init();
        tma_prefetch_desc(&tmap_x);
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), 512u);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_smem;

    if (warp == 0) {
        // ===================== TMA producer: one (tap, channel chunk) per stage =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                const int img = tile / p.tiles_per_img;
                const int trem = tile - img * p.tiles_per_img;
                const int ty = trem / p.tiles_x, tx = trem - ty * p.tiles_x;
                const int y0 = ty * p.BH, x0 = tx * p.BW;
                int ch = 0, kh = 0, kw = 0;
                for (int it = 0; it < nit; ++it) {
                    mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
                    const uint32_t full = smem_u32(&bar_full[s]);
                    mbar_arrive_expect_tx(full, (uint32_t)(p.x_bytes + p.w_bytes));
                    const uint32_t sa = smem_base + (uint32_t)s * (uint32_t)p.stage_bytes;
                    tma_load_4d(sa, &tmap_x, full, ch * p.kc, x0 + kw - p.pad_l, y0 + kh - p.pad_t, img);
                    bulk_load(sa + 2u * (uint32_t)p.x_bytes, p.wp + (size_t)it * 128 * p.kc, (uint32_t)p.w_bytes, full);
                    if (++ch == p.nchunks) { ch = 0; if (++kw == p.KW) { kw = 0; ++kh; } }
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(128, 256, 0, 0);
            const uint32_t sbo = 8u * (uint32_t)p.span;
            const int ksteps = p.kc >> 3;
            int s = 0;
            uint32_t ph = 0;
            int tcount = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tcount) {
                const int ab = tcount & 1;
                mbar_wait(smem_u32(&bar_tempty[ab]), (uint32_t)(((tcount >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t td = tmem_d + (uint32_t)(ab * 256);
                uint32_t accumulate = 0;
                for (int it = 0; it < nit; ++it) {
                    mbar_wait(smem_u32(&bar_conv[s]), ph);
                    tc_fence_after();
                    const uint32_t sx = smem_base + (uint32_t)s * (uint32_t)p.stage_bytes;
                    const uint32_t sw = sx + 2u * (uint32_t)p.x_bytes;
                    for (int k = 0; k < ksteps; ++k) {
                        const uint32_t ko = (uint32_t)k * 32u;
                        const uint64_t da = make_smem_desc(sw + ko, 16, sbo, p.layout);
                        const uint64_t dbh = make_smem_desc(sx + ko, 16, sbo, p.layout);
                        const uint64_t dbl = make_smem_desc(sx + (uint32_t)p.x_bytes + ko, 16, sbo, p.layout);
                        umma_tf32(td, da, dbh, idesc, accumulate);
                        umma_tf32(td, da, dbl, idesc, 1u);
                        accumulate = 1u;
                    }
                    umma_commit(smem_u32(&bar_empty[s]));
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
                umma_commit(smem_u32(&bar_tfull[ab]));
            }
        }
    } else if (warp < 10) {
        // ===================== epilogue (warps 2-9) =====================
        // accumulator lane = output channel (quadrants 0,1: hi rows c = lane index; quadrants 2,3: lo rows of channel
        // c = lane index - 64), column = pixel of the tile.  The two warps of a quadrant take one 128-column half each.
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const bool is_lo = q >= 2;
        const int c = (q & 1) * 32 + lane;                    // output channel of this thread
        const bool c_ok = c < p.Cout;
        float* const ex = reinterpret_cast<float*>(smem_al + p.ex_off) + c * kTtExPitch;
        const int r = p.d2s_r;
        const int Cd = p.Cout / (r * r);
        const int g = r > 1 ? c / Cd : 0, cc = r > 1 ? c - g * Cd : c;
        const int di = r > 1 ? g / r : 0, dj = r > 1 ? g - di * r : 0;
        const float bias = bias_s[c & 63];
        int tcount = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tcount) {
            const int ab = tcount & 1;
            const int img = tile / p.tiles_per_img;
            const int trem = tile - img * p.tiles_per_img;
            const int ty = trem / p.tiles_x, tx = trem - ty * p.tiles_x;
            mbar_wait(smem_u32(&bar_tfull[ab]), (uint32_t)((tcount >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * 256 + half * 128);
            if (is_lo) {
                // publish the lo rows, then release the accumulator
                for (int c0 = 0; c0 < 128; c0 += 16) {
                    float v[16];
                    tmem_ld16(taddr + (uint32_t)c0, v);
                    if (c_ok) {
                        float* dst = ex + half * 128 + c0;
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                }
                tc_fence_before();
                mbar_arrive_warp(smem_u32(&bar_tempty[ab]));
                asm volatile("bar.sync 1, 256;" ::: "memory");          // lo rows of this tile are in `ex`
                asm volatile("bar.sync 2, 256;" ::: "memory");          // ... and have been consumed
            } else {
                asm volatile("bar.sync 1, 256;" ::: "memory");
                for (int c0 = 0; c0 < 128; c0 += 16) {
                    float v[16];
                    tmem_ld16(taddr + (uint32_t)c0, v);
                    if (c_ok) {
                        const int px = half * 128 + c0;                  // first of 16 consecutive pixels of one image row
                        const int ry = px / p.BW, rx = px - ry * p.BW;
                        const int oy = ty * p.BH + ry, ox = tx * p.BW + rx;
                        const float* lo = ex + px;
                        const int64_t pix = ((int64_t)img * p.H + oy) * p.W + ox;
                        const float* resp = p.res ? p.res + pix * p.res_ld + c : nullptr;
                        float* yp;
                        int64_t ystep;
                        if (r == 1) {
                            yp = p.y + pix * p.y_ld + c;
                            ystep = p.y_ld;
                        } else {
                            const int64_t hp = ((int64_t)img * p.H * r + (int64_t)oy * r + di) * ((int64_t)p.W * r) + (int64_t)ox * r + dj;
                            yp = p.y + hp * p.y_ld + cc;
                            ystep = (int64_t)r * p.y_ld;
                        }
                        float lv[16];
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 t = *reinterpret_cast<const float4*>(lo + j);
                            lv[j] = t.x; lv[j + 1] = t.y; lv[j + 2] = t.z; lv[j + 3] = t.w;
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float o = v[j] + lv[j] + bias;
                            if (resp) o += __ldg(resp + (int64_t)j * p.res_ld);
                            o = apply_act(o, p.act);
                            float* d = yp + j * ystep;
                            if (p.beta) o += *d;
                            *d = o;
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive_warp(smem_u32(&bar_tempty[ab]));
                asm volatile("bar.sync 2, 256;" ::: "memory");
            }
        }
    } else {
        // ===================== hi/lo splitter of the pixel tile (warps 10-17) =====================
        const int et = threadIdx.x - 320;          // 0..255
        int s = 0;
        uint32_t ph = 0;
        const int units = p.x_bytes >> 4;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            for (int it = 0; it < nit; ++it) {
                mbar_wait(smem_u32(&bar_full[s]), ph);
                uint8_t* x_hi = smem_al + (size_t)s * p.stage_bytes;
                uint8_t* x_lo = x_hi + p.x_bytes;
#pragma unroll 2
                for (int u = et; u < units; u += 256) {
                    const float4 v = *reinterpret_cast<const float4*>(x_hi + u * 16);
                    float4 h, l;
                    h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
                    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
                    *reinterpret_cast<float4*>(x_hi + u * 16) = h;
                    *reinterpret_cast<float4*>(x_lo + u * 16) = l;
                }
                fence_proxy_async_smem();
                mbar_arrive_warp(smem_u32(&bar_conv[s]));
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, 512u);
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
static inline int t_kc(int Cin) { return Cin % 16 == 0 ? 16 : 8; }

bool fwd_t_channels_ok(int Cin, int Cout, int KH, int KW, int math_mode) {
    static const bool disabled = [] { const char* e = getenv("DL4DS_TC_NO_T"); return e && e[0] == '1'; }();
    if (disabled || math_mode != DL4DS_MATH_TF32X3) return false;
    return Cin % 8 == 0 && Cout % 8 == 0 && Cout <= 64 && KH * KW <= 25;
}

int64_t fwd_t_pack_floats(int taps, int Cin) {
    const int kc = t_kc(Cin);
    return (int64_t)taps * ((Cin + kc - 1) / kc) * 128 * kc;
}

int conv2d_pack_t(const float* w, int wmode, int KH, int KW, int Cin, int Cout, float* dst, cudaStream_t st) {
    const int kc = t_kc(Cin);
    const int nchunks = (Cin + kc - 1) / kc;
    const int64_t units = fwd_t_pack_floats(KH * KW, Cin) / 4;
    const int blocks = (int)((units + 255) / 256 > 1184 ? 1184 : (units + 255) / 256);
    pack_weights_t_kernel<<<blocks, 256, 0, st>>>(w, dst, KH * KW, Cin, Cout, kc, nchunks, wmode);
    return check_launch("pack_weights_t_kernel");
}

static bool t_tile_geometry(int H, int W, int* BW, int* BH) {
    int bw;
    if (W >= 256) {
        if (W % 256) return false;
        bw = 256;
    } else {
        if (W < 16 || 256 % W) return false;
        bw = W;
    }
    const int bh = 256 / bw;
    if (H % bh) return false;
    *BW = bw;
    *BH = bh;
    return true;
}

// wt: the T image produced by conv2d_pack_t for (a.w, a.wmode).  DL4DS_E_UNSUPPORTED outside the domain.
int conv2d_fwd_t(const ConvArgs& a, const float* wt, cudaStream_t st) {
    if (!fwd_t_channels_ok(a.Cin, a.Cout, a.KH, a.KW, DL4DS_MATH_TF32X3)) return DL4DS_E_UNSUPPORTED;
    if (a.stride != 1 || a.up != 1 || a.Ho != a.H || a.Wo != a.W) return DL4DS_E_UNSUPPORTED;
    if (a.d2s_r != 1 && (a.d2s_r != 2 || a.Cout % 4)) return DL4DS_E_UNSUPPORTED;
    if (a.x_ld % 4 || (reinterpret_cast<uintptr_t>(a.x) & 15)) return DL4DS_E_UNSUPPORTED;
    TcTParams p;
    if (!t_tile_geometry(a.H, a.W, &p.BW, &p.BH)) return DL4DS_E_UNSUPPORTED;
    const int kc = t_kc(a.Cin);
    p.wp = wt; p.bias = a.bias; p.res = a.res; p.y = a.y; p.res_ld = a.res_ld; p.y_ld = a.y_ld;
    p.H = a.H; p.W = a.W; p.Cin = a.Cin; p.Cout = a.Cout;
    p.KW = a.KW; p.ntaps = a.KH * a.KW; p.pad_t = a.pad_t; p.pad_l = a.pad_l;
    p.tiles_x = a.W / p.BW;
    p.tiles_per_img = p.tiles_x * (a.H / p.BH);
    p.ntiles = a.N * p.tiles_per_img;
    p.kc = kc; p.span = kc * 4; p.nchunks = (a.Cin + kc - 1) / kc;
    p.layout = kc == 16 ? kLayoutSw64 : kLayoutSw32;
    p.act = a.act; p.d2s_r = a.d2s_r; p.beta = a.beta;
    p.x_bytes = 256 * p.span;
    p.w_bytes = 128 * p.span;
    p.stage_bytes = 2 * p.x_bytes + p.w_bytes;
    const int ex_bytes = 64 * kTtExPitch * 4;
    int stages = (int)((218 * 1024 - ex_bytes) / p.stage_bytes);
    if (stages > kTtMaxStages) stages = kTtMaxStages;
    if (stages < 2) return DL4DS_E_UNSUPPORTED;
    p.stages = stages;
    p.ex_off = stages * p.stage_bytes;
    const size_t smem = (size_t)p.ex_off + ex_bytes + 1024;
    const CUtensorMap* tm = get_tensor_map_nhwc(a.x, a.x_ld, a.N, a.H, a.W, a.Cin, kc, p.BW, p.BH,
                                                kc == 16 ? (int)CU_TENSOR_MAP_SWIZZLE_64B : (int)CU_TENSOR_MAP_SWIZZLE_32B);
    if (!tm) return DL4DS_E_CUDA;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(conv_tc_fwdT_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(222 * 1024));
        attr = true;
    }
    const int grid = p.ntiles < kNumSMs ? p.ntiles : kNumSMs;
    conv_tc_fwdT_kernel<<<grid, kTtThreads, smem, st>>>(*tm, p);
    g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    return check_launch("conv_tc_fwdT_kernel");
}

}  // namespace dl4ds
