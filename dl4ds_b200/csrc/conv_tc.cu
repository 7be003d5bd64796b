// tcgen05 (5th-gen tensor core) implicit-GEMM convolution kernels for sm_100a.
//
// Forward / input-gradient (stride-1 'same'-grid convolutions -- blocks.py:49-61,208,299,414-416;
// sp_postups.py:134,156 -- and their dgrads, which are the same op with flipped/transposed weights):
//   GEMM M = 128 output pixels (a BH x BW patch of one image), N = Cout (padded to 16), K = taps x Cin.
//   A (activations): one TMA 4-D box {KC ch, BW, BH, 1} per (tap, channel chunk), shifted by the
//     tap offset, out-of-bounds rows/cols/channels zero-filled by TMA = the convolution's zero padding;
//     lands K-major in the hardware 32/64/128-byte swizzle.
//   B (weights): pre-packed once per optimizer step by pack_weights_kernel into the exact swizzled
//     shared-memory image ([tap][chunk][Npad][KC]), fetched with one cp.async.bulk per stage.
//   D: fp32 accumulator in TMEM (Npad columns x 128 lanes), tcgen05.mma kind::tf32 issued by one thread.
//   Epilogue (4 warps): tcgen05.ld -> +bias (+residual) -> activation -> vectorised NHWC store, or the
//     depth_to_space (tf.nn.depth_to_space, DCR order, blocks.py:427) permuted store.
// Math modes: DL4DS_MATH_TF32 feeds raw fp32 (the tensor core reads the top 19 bits);
// DL4DS_MATH_TF32X3 splits both operands into tf32 hi + lo and issues hi*lo + lo*hi + hi*hi
// (fp32 accumulate) -- error ~2^-21 per product, i.e. fp32-level parity.  The activation split runs
// in-kernel on the (otherwise idle) epilogue warps; the weight split is done by the pack kernel.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2-5 = operand splitter (x3 mode) during the main loop, then epilogue.
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <tuple>

#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace dl4ds {

using namespace tc;

// -------------------------------------------------------------------------------------------------
// host: tensor-map cache
// -------------------------------------------------------------------------------------------------
namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

const CUtensorMap* get_tensor_map_nhwc(const float* base, int ld, int N, int H, int W, int C,
                                       int box_c, int box_w, int box_h, int swizzle) {
    typedef std::tuple<const void*, int, int, int, int, int, int, int, int, int> Key;
    static std::map<Key, CUtensorMap*> cache;
    static std::mutex mu;
    Key key(base, ld, N, H, W, C, box_c, box_w, box_h, swizzle);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    EncodeTiledFn enc = encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return nullptr;
    }
    CUtensorMap* m = new CUtensorMap;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)W * ld * 4, (cuuint64_t)H * W * ld * 4};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swizzle,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for C=%d W=%d H=%d N=%d ld=%d box=%d,%d,%d", (int)r, C, W, H, N,
                  ld, box_c, box_w, box_h);
        delete m;
        return nullptr;
    }
    if (cache.size() > 4096) {          // unbounded growth guard (addresses are stable under CUDA graphs)
        for (auto& kv : cache) delete kv.second;
        cache.clear();
    }
    cache[key] = m;
    return m;
}

}  // namespace tc

// -------------------------------------------------------------------------------------------------
// weight packing: Keras-layout weights -> [tap][chunk][Npad][KC] swizzled smem images (hi and lo)
// -------------------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo,
                                    int taps, int Cin, int Cout, int Npad, int kc, int nchunks, int wmode,
                                    int x3) {
    const int upr = kc / 4;                                  // 16-byte units per row
    const int64_t total = (int64_t)taps * nchunks * Npad * upr;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int u = (int)(idx % upr);
        const int n = (int)((idx / upr) % Npad);
        const int blk = (int)(idx / ((int64_t)upr * Npad));
        const int tap = blk / nchunks, ch = blk - tap * nchunks;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = ch * kc + u * 4 + j;
            float x = 0.0f;
            if (n < Cout && c < Cin) {
                if (wmode == DL4DS_W_HWIO)
                    x = __ldg(w + ((int64_t)tap * Cin + c) * Cout + n);
                else
                    x = __ldg(w + ((int64_t)(taps - 1 - tap) * Cout + n) * Cin + c);
            }
            v[j] = x;
        }
        const int us = swizzle_unit(u, n, kc * 4);
        const int64_t dst = ((int64_t)blk * Npad + n) * kc + us * 4;
        if (x3) {
            float h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                h[j] = tf32_rna(v[j]);
                l[j] = tf32_rna(v[j] - h[j]);      // on the tf32 grid: the tensor core's truncation is then exact
            }
            *reinterpret_cast<float4*>(hi + dst) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(lo + dst) = make_float4(l[0], l[1], l[2], l[3]);
        } else {
            *reinterpret_cast<float4*>(hi + dst) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// -------------------------------------------------------------------------------------------------
// 3-term fp16 mode (DL4DS_MATH_F16X3): the weight image [tap][chunk][Npad][KC16] as fp16 hi / lo halves of w * s_w,
// s_w = the power of two that brings max |w| of the layer into [2^13, 2^14) (fp16 keeps 11 significant bits like tf32
// but a 5-bit exponent).  The scale pair {s_w, 1 / s_w} sits behind the two images; the convolution's epilogue
// multiplies the accumulator by 1 / (s_w * s_tile).  Geometry: tc::pick_chunk16.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) weight_absmax_kernel(const float* __restrict__ w, int64_t n, float* __restrict__ scale_out) {
    __shared__ float red[8];
    float m = 0.0f;
    for (int64_t i = threadIdx.x; i < n; i += 256) m = fmaxf(m, fabsf(__ldg(w + i)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int j = 1; j < 8; ++j) m = fmaxf(m, red[j]);
        const float s = pow2_scale_for(m);
        scale_out[0] = s;
        scale_out[1] = 1.0f / s;
    }
}

// one 16-byte unit (8 halfs) of the hi and lo fp16 images; li = unit index inside the layer's fp16 image
__device__ __forceinline__ void pack_f16_unit(const float* __restrict__ w, __half* __restrict__ hi, __half* __restrict__ lo,
                                              const float* __restrict__ scale, int64_t li, int taps, int Cin, int Cout,
                                              int Npad, int kc16, int nchunks16, int wmode) {
    const int upr = kc16 / 8;
    const int u = (int)(li % upr);
    const int nn = (int)((li / upr) % Npad);
    const int blk = (int)(li / ((int64_t)upr * Npad));
    const int tap = blk / nchunks16, ch = blk - tap * nchunks16;
    const float s = __ldg(scale);
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = ch * kc16 + u * 8 + j;
        float x = 0.0f;
        if (nn < Cout && c < Cin) {
            if (wmode == DL4DS_W_HWIO)
                x = __ldg(w + ((int64_t)tap * Cin + c) * Cout + nn);
            else
                x = __ldg(w + ((int64_t)(taps - 1 - tap) * Cout + nn) * Cin + c);
        }
        x *= s;
        h[j] = __float2half_rn(x);
        l[j] = __float2half_rn(x - __half2float(h[j]));
    }
    const int us = swizzle_unit(u, nn, kc16 * 2);
    const int64_t dst = ((int64_t)blk * Npad + nn) * kc16 + us * 8;
    *reinterpret_cast<uint4*>(hi + dst) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(lo + dst) = *reinterpret_cast<const uint4*>(l);
}

struct F16Image {       // where the fp16 part of a DL4DS_MATH_F16X3 workspace lives (behind the two tf32 images)
    int kc16, nchunks16;
    int64_t n16;        // halfs per image
    __half* hi; __half* lo; float* scale;
};
__host__ __device__ inline F16Image f16_image(float* tf32_hi, int64_t n_tf32, int taps, int Cin, int Npad) {
    F16Image f;
    const int cp = (Cin + 15) / 16 * 16;
    f.kc16 = (cp % 64 == 0) ? 64 : ((cp % 32 == 0) ? 32 : 16);
    f.nchunks16 = cp / f.kc16;
    f.n16 = (int64_t)taps * f.nchunks16 * Npad * f.kc16;
    f.hi = reinterpret_cast<__half*>(tf32_hi + 2 * n_tf32);
    f.lo = f.hi + f.n16;
    f.scale = reinterpret_cast<float*>(f.lo + f.n16);
    return f;
}

__global__ void pack_weights_f16_kernel(const float* __restrict__ w, float* __restrict__ tf32_hi, int64_t n_tf32, int taps,
                                        int Cin, int Cout, int Npad, int wmode) {
    const F16Image f = f16_image(tf32_hi, n_tf32, taps, Cin, Npad);
    const int64_t total = f.n16 / 8;
    for (int64_t li = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; li < total; li += (int64_t)gridDim.x * blockDim.x)
        pack_f16_unit(w, f.hi, f.lo, f.scale, li, taps, Cin, Cout, Npad, f.kc16, f.nchunks16, wmode);
}

// Every (layer, pass) weight image of a model in ONE launch: a table of descriptors in device memory, a flat unit
// index space (16-byte units of all images), each unit located by a binary search over the descriptors' first units.
// Replaces ~40 three-microsecond pack launches per optimizer step (and the side-stream fork / join around them).
struct PackDesc {
    const float* w; float* hi; float* lo;
    int taps, Cin, Cout, Npad, kc, nchunks, wmode, x3;
    long long unit_begin;
};
static_assert(sizeof(PackDesc) == 64, "PackDesc is a 64-byte record (dl4ds_conv2d_pack_desc writes it, Python fills unit_begin)");

// one block per descriptor of a DL4DS_MATH_F16X3 pack table (x3 == 2): the layer's weight scale
__global__ void __launch_bounds__(256) weight_absmax_multi_kernel(const PackDesc* __restrict__ descs) {
    __shared__ float red[8];
    const PackDesc d = descs[blockIdx.x];
    if (d.x3 != 2) return;
    const int64_t n = (int64_t)d.taps * d.Cin * d.Cout;
    float m = 0.0f;
    for (int64_t i = threadIdx.x; i < n; i += 256) m = fmaxf(m, fabsf(__ldg(d.w + i)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int j = 1; j < 8; ++j) m = fmaxf(m, red[j]);
        const F16Image f = f16_image(d.hi, (int64_t)d.taps * d.nchunks * d.Npad * d.kc, d.taps, d.Cin, d.Npad);
        const float s = pow2_scale_for(m);
        f.scale[0] = s;
        f.scale[1] = 1.0f / s;
    }
}

__global__ void pack_weights_multi_kernel(const PackDesc* __restrict__ descs, int n, long long total_units) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total_units;
         idx += (long long)gridDim.x * blockDim.x) {
        int lo_i = 0, hi_i = n - 1;
        while (lo_i < hi_i) {                         // last descriptor whose unit_begin <= idx
            const int mid = (lo_i + hi_i + 1) >> 1;
            if (descs[mid].unit_begin <= idx) lo_i = mid; else hi_i = mid - 1;
        }
        const PackDesc d = descs[lo_i];
        long long li = idx - d.unit_begin;
        const long long n_tf32 = (long long)d.taps * d.nchunks * d.Npad * d.kc;
        if (li >= n_tf32 / 4) {          // x3 == 2: the fp16 images follow the two tf32 ones
            const F16Image f = f16_image(d.hi, n_tf32, d.taps, d.Cin, d.Npad);
            pack_f16_unit(d.w, f.hi, f.lo, f.scale, li - n_tf32 / 4, d.taps, d.Cin, d.Cout, d.Npad, f.kc16, f.nchunks16,
                          d.wmode);
            continue;
        }
        const int upr = d.kc / 4;
        const int u = (int)(li % upr);
        const int nn = (int)((li / upr) % d.Npad);
        const int blk = (int)(li / ((long long)upr * d.Npad));
        const int tap = blk / d.nchunks, ch = blk - tap * d.nchunks;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = ch * d.kc + u * 4 + j;
            float x = 0.0f;
            if (nn < d.Cout && c < d.Cin) {
                if (d.wmode == DL4DS_W_HWIO)
                    x = __ldg(d.w + ((int64_t)tap * d.Cin + c) * d.Cout + nn);
                else
                    x = __ldg(d.w + ((int64_t)(d.taps - 1 - tap) * d.Cout + nn) * d.Cin + c);
            }
            v[j] = x;
        }
        const int us = swizzle_unit(u, nn, d.kc * 4);
        const int64_t dst = ((int64_t)blk * d.Npad + nn) * d.kc + us * 4;
        if (d.x3) {
            float h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                h[j] = tf32_rna(v[j]);
                l[j] = tf32_rna(v[j] - h[j]);
            }
            *reinterpret_cast<float4*>(d.hi + dst) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(d.lo + dst) = make_float4(l[0], l[1], l[2], l[3]);
        } else {
            *reinterpret_cast<float4*>(d.hi + dst) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// -------------------------------------------------------------------------------------------------
// forward / dgrad kernel
// -------------------------------------------------------------------------------------------------
struct TcFwdParams {
    const float* wp_hi;
    const float* wp_lo;
    const float* bias;
    const float* res;
    float* y;
    int res_ld, y_ld;
    int H, W, Cin, Cout, Npad;
    int KW, ntaps, pad_t, pad_l;
    int BW, BH, tiles_x, tiles_per_img;
    int kc, span, nchunks;
    uint32_t layout;
    int act, d2s_r, beta;
    int stages, stage_bytes, a_bytes, b_bytes, tmem_cols, ntiles;
    int group, ngroups;     // kernel-tap / channel-chunk iterations per smem stage, stages per tile
    int stackn;             // x3, Npad <= 64: B = [W_hi ; W_lo] stacked on N -> 2 MMAs per K-step (see the issuer)
    int acc_stride;         // TMEM columns per accumulator buffer: Npad, or 2*Npad when stackn
    int lo_tmem;            // x3 + stackn: the lo halves of the activation tile go to TMEM (A operand of the second MMA
                            // from TMEM), the raw tile in smem is the hi operand: no A_lo write / read in shared memory
    int a_tmem;             // x3: split activations go to TMEM (A operand from TMEM), not back to smem
    int tmem_a_off;         // first TMEM column of the A ring (stage s, iteration j: + (s*group+j)*2*kc)
};

constexpr int kTcThreads = 320;       // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int kTcThreadsX3 = 448;     // + warps 10-13: tf32 hi/lo splitter
constexpr int kMaxStages = 16;

// Persistent: grid = min(#tiles, #SMs); every CTA walks tiles blockIdx.x, +gridDim.x, ...  The smem
// stage ring runs continuously across tiles and the accumulator is double-buffered in TMEM
// (2 x Npad columns), so the epilogue of tile i overlaps the TMA / MMA work of tile i+1.
// AT: compile the two experimental A-operand-in-TMEM paths (p.a_tmem / p.lo_tmem) -- kept out of the default
// instantiation, where their splitter code cost the hot kernel 84 bytes of register spills
template <bool X3, bool AT>
__global__ void __launch_bounds__(X3 ? kTcThreadsX3 : kTcThreads, 2)
conv_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmap_x, const TcFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[kMaxStages];
    __shared__ __align__(8) uint64_t bar_conv[kMaxStages];
    __shared__ __align__(8) uint64_t bar_empty[kMaxStages];
    __shared__ __align__(8) uint64_t bar_tfull[2];
    __shared__ __align__(8) uint64_t bar_tempty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float bias_s[256];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int nit = p.ntaps * p.nchunks;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) bias_s[i] = (p.bias && i < p.Cout) ? __ldg(p.bias + i) : 0.0f;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(&bar_full[s]), 1);
            mbar_init(smem_u32(&bar_conv[s]), 4);
            mbar_init(smem_u32(&bar_empty[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&bar_tfull[b]), 1);
            mbar_init(smem_u32(&bar_tempty[b]), 8);
        }
        fence_barrier_init();
        tma_prefetch_desc(&tmap_x);
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_smem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const uint32_t it_bytes = (uint32_t)(p.a_bytes + p.b_bytes * (X3 ? 2 : 1));
            const uint32_t a_lo_off = (uint32_t)(p.group * p.a_bytes);
            const uint32_t b_off = a_lo_off * ((X3 && !p.a_tmem && !p.lo_tmem) ? 2u : 1u);
            const uint32_t b_lo_off = (uint32_t)(p.group * p.b_bytes);
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                const int img = tile / p.tiles_per_img;
                const int trem = tile - img * p.tiles_per_img;
                const int ty = trem / p.tiles_x, tx = trem - ty * p.tiles_x;
                const int y0 = ty * p.BH, x0 = tx * p.BW;
                int ch = 0, kh = 0, kw = 0;
                for (int g = 0; g < p.ngroups; ++g) {
                    const int it0 = g * p.group;
                    const int n = min(p.group, nit - it0);
                    mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
                    const uint32_t full = smem_u32(&bar_full[s]);
                    mbar_arrive_expect_tx(full, (uint32_t)n * it_bytes);
                    const uint32_t sa = smem_base + (uint32_t)s * (uint32_t)p.stage_bytes;
                    for (int j = 0; j < n; ++j) {
                        tma_load_4d(sa + (uint32_t)(j * p.a_bytes), &tmap_x, full, ch * p.kc, x0 + kw - p.pad_l,
                                    y0 + kh - p.pad_t, img);
                        if (++ch == p.nchunks) { ch = 0; if (++kw == p.KW) { kw = 0; ++kh; } }
                    }
                    const size_t woff = (size_t)it0 * p.Npad * p.kc;
                    if (X3 && p.stackn) {
                        // hi and lo tiles of one iteration adjacent in shared memory: one B operand of 2*Npad rows
                        for (int j = 0; j < n; ++j) {
                            const size_t wj = woff + (size_t)j * p.Npad * p.kc;
                            bulk_load(sa + b_off + (uint32_t)(2 * j * p.b_bytes), p.wp_hi + wj, (uint32_t)p.b_bytes, full);
                            bulk_load(sa + b_off + (uint32_t)((2 * j + 1) * p.b_bytes), p.wp_lo + wj, (uint32_t)p.b_bytes, full);
                        }
                    } else {
                        bulk_load(sa + b_off, p.wp_hi + woff, (uint32_t)(n * p.b_bytes), full);
                        if (X3) bulk_load(sa + b_off + b_lo_off, p.wp_lo + woff, (uint32_t)(n * p.b_bytes), full);
                    }
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(128, p.Npad, 0, 0);
            const uint32_t idesc2 = make_idesc_tf32(128, 2 * p.Npad, 0, 0);
            const uint32_t sbo = 8u * (uint32_t)p.span;
            int s = 0;
            uint32_t ph = 0;
            int tcount = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tcount) {
                const int ab = tcount & 1;
                mbar_wait(smem_u32(&bar_tempty[ab]), (uint32_t)(((tcount >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t td = tmem_d + (uint32_t)(ab * p.acc_stride);
                uint32_t accumulate = 0;
                int ch = 0;
                const uint32_t a_lo_off = (uint32_t)(p.group * p.a_bytes);
                const uint32_t b_off = a_lo_off * ((X3 && !p.a_tmem && !p.lo_tmem) ? 2u : 1u);
                const uint32_t b_lo_off = (uint32_t)(p.group * p.b_bytes);
                for (int g = 0; g < p.ngroups; ++g) {
                    const int n = min(p.group, nit - g * p.group);
                    mbar_wait(smem_u32(X3 ? &bar_conv[s] : &bar_full[s]), ph);
                    tc_fence_after();
                    const uint32_t st0 = smem_base + (uint32_t)s * (uint32_t)p.stage_bytes;
                    for (int j = 0; j < n; ++j) {
                        int ksteps = (p.Cin - ch * p.kc);
                        ksteps = (ksteps > p.kc ? p.kc : ksteps) >> 3;
                        const uint32_t sa = st0 + (uint32_t)(j * p.a_bytes);
                        const uint32_t sb = st0 + b_off + (uint32_t)(j * p.b_bytes * (p.stackn ? 2 : 1));
                        const uint32_t ta = tmem_d + (uint32_t)(p.tmem_a_off + (s * p.group + j) * 2 * p.kc);
                        for (int k = 0; k < ksteps; ++k) {
                            const uint32_t ko = (uint32_t)k * 32u;
                            const uint64_t da = make_smem_desc(sa + ko, 16, sbo, p.layout);
                            const uint64_t db = make_smem_desc(sb + ko, 16, sbo, p.layout);
                            if (X3 && AT && p.a_tmem) {
                                const uint64_t dbl = make_smem_desc(sb + b_lo_off + ko, 16, sbo, p.layout);
                                const uint32_t ah = ta + (uint32_t)(k * 8), al = ah + (uint32_t)p.kc;
                                umma_tf32_ts(td, al, db, idesc, accumulate);
                                umma_tf32_ts(td, ah, dbl, idesc, 1u);
                                umma_tf32_ts(td, ah, db, idesc, 1u);
                            } else if (X3 && AT && p.lo_tmem) {
                                // as the stacked-N scheme below, with A_lo read from TMEM (lane = pixel row, one
                                // column per channel of the chunk) instead of shared memory
                                const uint32_t tl = tmem_d + (uint32_t)(p.tmem_a_off + (s * p.group + j) * p.kc + k * 8);
                                umma_tf32(td, da, db, idesc2, accumulate);
                                umma_tf32_ts(td, tl, db, idesc, 1u);
                            } else if (X3 && p.stackn) {
                                // A_hi x [W_hi ; W_lo] -> columns [0,Npad) and [Npad,2Npad); A_lo x W_hi -> columns
                                // [0,Npad).  One MMA costs ~119 cycles for any N <= 128 (scratch/umma_rate.cu), so
                                // 2 instead of 3 per K-step; the epilogue adds the two column groups.
                                const uint64_t dal = make_smem_desc(sa + a_lo_off + ko, 16, sbo, p.layout);
                                umma_tf32(td, da, db, idesc2, accumulate);
                                umma_tf32(td, dal, db, idesc, 1u);
                            } else if (X3) {
                                const uint64_t dal = make_smem_desc(sa + a_lo_off + ko, 16, sbo, p.layout);
                                const uint64_t dbl = make_smem_desc(sb + b_lo_off + ko, 16, sbo, p.layout);
                                umma_tf32(td, dal, db, idesc, accumulate);
                                umma_tf32(td, da, dbl, idesc, 1u);
                                umma_tf32(td, da, db, idesc, 1u);
                            } else {
                                umma_tf32(td, da, db, idesc, accumulate);
                            }
                            accumulate = 1u;
                        }
                        if (++ch == p.nchunks) ch = 0;
                    }
                    umma_commit(smem_u32(&bar_empty[s]));
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
                umma_commit(smem_u32(&bar_tfull[ab]));
            }
        }
    } else if (warp < 10) {
        // ===================== epilogue (warps 2-9) =====================
        // TMEM lane quadrant q = warp % 4 (hardware rule); the two warps of a quadrant take alternate
        // 16-column blocks.  Thread = one output pixel (TMEM lane), 16 consecutive channels per block.
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const int ry = row / p.BW, rx = row - ry * p.BW;
        const int r = p.d2s_r;
        const int Cd = p.Cout / (r * r);
        int tcount = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tcount) {
            const int ab = tcount & 1;
            const int img = tile / p.tiles_per_img;
            const int trem = tile - img * p.tiles_per_img;
            const int ty = trem / p.tiles_x, tx = trem - ty * p.tiles_x;
            const int oy = ty * p.BH + ry, ox = tx * p.BW + rx;
            const int64_t pix = ((int64_t)img * p.H + oy) * p.W + ox;
            const float* __restrict__ resp = p.res ? p.res + pix * p.res_ld : nullptr;
            float* __restrict__ yp = p.y + pix * p.y_ld;
            // depth_to_space: HR pixel (oy*r+di, ox*r+dj) receives channels [g*Cd, (g+1)*Cd), g = di*r+dj
            const int64_t hr_row0 = ((int64_t)img * p.H * r + (int64_t)oy * r) * ((int64_t)p.W * r) + (int64_t)ox * r;
            mbar_wait(smem_u32(&bar_tfull[ab]), (uint32_t)((tcount >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * p.acc_stride);
            for (int c0 = half * 16; c0 < p.Npad; c0 += 32) {
                float v[16];
                tmem_ld16(taddr + (uint32_t)c0, v);
                if (X3 && p.stackn) {
                    float v2[16];
                    tmem_ld16(taddr + (uint32_t)(p.Npad + c0), v2);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += v2[j];
                }
                if (c0 >= p.Cout) continue;
                float4 rs[4];
                if (resp) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        rs[j] = (c0 + 4 * j < p.Cout) ? __ldg(reinterpret_cast<const float4*>(resp + c0) + j)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                int g = 0, cg = c0;
                if (r > 1) { g = c0 / Cd; cg = c0 - g * Cd; }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int co = c0 + 4 * j;
                    if (co >= p.Cout) break;
                    const float4 b = *reinterpret_cast<const float4*>(&bias_s[co]);
                    float4 o = make_float4(v[4 * j] + b.x, v[4 * j + 1] + b.y, v[4 * j + 2] + b.z, v[4 * j + 3] + b.w);
                    if (resp) { o.x += rs[j].x; o.y += rs[j].y; o.z += rs[j].z; o.w += rs[j].w; }
                    o.x = apply_act(o.x, p.act); o.y = apply_act(o.y, p.act);
                    o.z = apply_act(o.z, p.act); o.w = apply_act(o.w, p.act);
                    if (r == 1) {
                        float4* dst = reinterpret_cast<float4*>(yp + co);
                        if (p.beta) {
                            const float4 old = *dst;
                            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                        }
                        *dst = o;
                    } else {
                        if (cg >= Cd) { cg -= Cd; ++g; }
                        const int di = g / r, dj = g - di * r;
                        const int64_t hp = hr_row0 + (int64_t)di * p.W * r + dj;
                        *reinterpret_cast<float4*>(p.y + hp * p.y_ld + cg) = o;
                        cg += 4;
                    }
                }
            }
            // accumulator buffer drained: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive_warp(smem_u32(&bar_tempty[ab]));
        }
    } else if (X3) {
        // ===================== operand splitter (warps 10-13, x3 mode) =====================
        const int et = threadIdx.x - 320;          // 0..127
        int s = 0;
        uint32_t ph = 0;
        uint8_t* const smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
        const int a_lo_off = p.group * p.a_bytes;
        int pending = -1;                          // TS mode: stage whose TMEM stores are issued but not yet published
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            for (int g = 0; g < p.ngroups; ++g) {
                const int n = min(p.group, nit - g * p.group);
                mbar_wait(smem_u32(&bar_full[s]), ph);
                uint8_t* a_hi = smem_al + (size_t)s * p.stage_bytes;
                uint8_t* a_lo = a_hi + a_lo_off;
                if (AT && p.lo_tmem) {
                    // thread = pixel row of the tile = TMEM lane: read the row's channels of every chunk of the stage
                    // through the TMA swizzle (a quarter-warp = 8 consecutive rows hits 8 distinct 16-byte slots),
                    // keep the raw tile untouched (it is the hi operand) and store lo = v - trunc(v) to TMEM.  The
                    // arrival for stage s is deferred until stage s+1's shared-memory reads are issued, which
                    // keeps the tcgen05.st latency off the critical path.
                    const int q = warp & 3;
                    const int row = q * 32 + lane;
                    const uint32_t trow = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)p.tmem_a_off;
                    const int noct = (n * p.kc) >> 3;                     // <= 4 (group * kc <= 32)
                    float4 v0[4], v1[4];
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        if (o < noct) {
                            const int e0 = o * 8, j = e0 / p.kc, c = e0 - j * p.kc;
                            const uint8_t* arow = a_hi + (size_t)j * p.a_bytes + (size_t)row * p.span;
                            v0[o] = *reinterpret_cast<const float4*>(arow + (swizzle_unit(c >> 2, row, p.span) << 4));
                            v1[o] = *reinterpret_cast<const float4*>(arow + (swizzle_unit((c >> 2) + 1, row, p.span) << 4));
                        }
                    }
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        if (o < noct) {
                            const int e0 = o * 8, j = e0 / p.kc, c = e0 - j * p.kc;
                            const float vv[8] = {v0[o].x, v0[o].y, v0[o].z, v0[o].w, v1[o].x, v1[o].y, v1[o].z, v1[o].w};
                            float l[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) l[e] = tf32_lo_of_trunc(vv[e]);
                            tmem_st8(trow + (uint32_t)((s * p.group + j) * p.kc + c), l);
                        }
                    }
                    tmem_wait_st();
                    tc_fence_before();
                    mbar_arrive_warp(smem_u32(&bar_conv[s]));
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                    continue;
                }
                if (AT && p.a_tmem) {
                    // thread = A row (TMEM lane): read the row's channels of the stage (<= 32 floats = 4 octets)
                    // through the TMA swizzle, then -- only now -- wait for the PREVIOUS stage's TMEM stores
                    // and publish that stage, then split and store this stage (asynchronously).  The deferred
                    // arrive keeps the tcgen05.st latency off the critical path.
                    const int q = warp & 3;
                    const int row = q * 32 + lane;
                    const uint32_t trow = tmem_d + ((uint32_t)(q * 32) << 16);
                    const int noct = (n * p.kc) >> 3;
                    float4 v0[4], v1[4];
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        if (o < noct) {
                            const int e0 = o * 8, j = e0 / p.kc, c = e0 - j * p.kc;
                            const uint8_t* arow = a_hi + (size_t)j * p.a_bytes + (size_t)row * p.span;
                            v0[o] = *reinterpret_cast<const float4*>(arow + (swizzle_unit(c >> 2, row, p.span) << 4));
                            v1[o] = *reinterpret_cast<const float4*>(arow + (swizzle_unit((c >> 2) + 1, row, p.span) << 4));
                        }
                    }
                    if (pending >= 0) {
                        tmem_wait_st();
                        tc_fence_before();
                        mbar_arrive_warp(smem_u32(&bar_conv[pending]));
                    }
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        if (o < noct) {
                            const int e0 = o * 8, j = e0 / p.kc, c = e0 - j * p.kc;
                            const uint32_t ta = trow + (uint32_t)(p.tmem_a_off + (s * p.group + j) * 2 * p.kc);
                            const float vv[8] = {v0[o].x, v0[o].y, v0[o].z, v0[o].w, v1[o].x, v1[o].y, v1[o].z, v1[o].w};
                            float h[8], l[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) tf32_split(vv[e], h[e], l[e]);
                            tmem_st8(ta + (uint32_t)c, h);
                            tmem_st8(ta + (uint32_t)(p.kc + c), l);
                        }
                    }
                    pending = s;
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                    continue;
                }
                const int units = n * (p.a_bytes >> 4);
#pragma unroll 2
                // The raw fp32 tile stays where TMA put it and serves as the hi operand: kind::tf32 reads the top 19
                // bits of each element, i.e. hi = trunc(v).  Only lo = v - trunc(v) (exact: <= 13 significant bits) is
                // written -- one 16-byte store per unit instead of two; these kernels are bound by shared-memory
                // bandwidth (~57 KB per 256 pixels and K slice, measured ~92 B/cycle/SM), not by the tensor pipe.
                for (int u = et; u < units; u += 128) {
                    const float4 v = *reinterpret_cast<const float4*>(a_hi + u * 16);
                    float4 l;
                    l.x = tf32_lo_of_trunc(v.x); l.y = tf32_lo_of_trunc(v.y);
                    l.z = tf32_lo_of_trunc(v.z); l.w = tf32_lo_of_trunc(v.w);
                    *reinterpret_cast<float4*>(a_lo + u * 16) = l;
                }
                fence_proxy_async_smem();
                mbar_arrive_warp(smem_u32(&bar_conv[s]));
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        }
        if (pending >= 0) {
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive_warp(smem_u32(&bar_conv[pending]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, (uint32_t)p.tmem_cols);
}

// -------------------------------------------------------------------------------------------------
// host dispatch
// -------------------------------------------------------------------------------------------------
std::atomic<long long> g_tc_launches{0};

static bool is_sm100() {
    static int v = -1;
    if (v < 0) v = dl4ds_device_is_sm100();
    return v == 1;
}

static bool tile_geometry(int H, int W, int tile_pix, int* BW, int* BH) {
    int bw;
    if (W >= tile_pix) {
        if (W % tile_pix) return false;
        bw = tile_pix;
    } else {
        if (W < 8 || tile_pix % W) return false;
        bw = W;
    }
    const int bh = tile_pix / bw;
    if (bh > 256 || H % bh) return false;
    *BW = bw;
    *BH = bh;
    return true;
}

// shape-only part of the eligibility test (what the workspace query can see)
static bool fwd_shape_supported(const ConvArgs& a, int math_mode) {
    if (math_mode != DL4DS_MATH_TF32 && math_mode != DL4DS_MATH_TF32X3) return false;
    if (!is_sm100()) return false;
    if (a.stride != 1 || a.up != 1 || a.Ho != a.H || a.Wo != a.W) return false;
    if (a.Cin % 8 || a.Cout % 8 || a.Cout > 256) return false;
    if (a.d2s_r != 1 && (a.d2s_r != 2 || (a.Cout / 4) % 4)) return false;
    if (a.KH * a.KW > 81) return false;
    int bw, bh;
    return tile_geometry(a.H, a.W, 128, &bw, &bh) || conv2d_fwd_halo_supported(a, math_mode);
}

static bool fwd_supported(const ConvArgs& a, int math_mode) {
    if (!fwd_shape_supported(a, math_mode)) return false;
    if (a.x_ld % 4 || (reinterpret_cast<uintptr_t>(a.x) & 15)) return false;
    if (a.y_ld % 4 || (reinterpret_cast<uintptr_t>(a.y) & 15)) return false;
    if (a.res && (a.res_ld % 4 || (reinterpret_cast<uintptr_t>(a.res) & 15))) return false;
    if (a.bias && (reinterpret_cast<uintptr_t>(a.bias) & 15)) return false;
    return true;
}

static int64_t pack_floats(int taps, int Cin, int Cout) {
    const Chunk c = pick_chunk(Cin);
    const int nchunks = (Cin + c.kc - 1) / c.kc;
    const int npad = (Cout + 15) / 16 * 16;
    return (int64_t)taps * nchunks * npad * c.kc;
}

int64_t conv2d_fwd_tc_workspace(const ConvArgs& a, int math_mode) {
    if (!fwd_shape_supported(a, math_mode == DL4DS_MATH_F16X3 ? DL4DS_MATH_TF32X3 : math_mode)) return 0;
    const int64_t n = pack_floats(a.KH * a.KW, a.Cin, a.Cout);
    if (math_mode == DL4DS_MATH_F16X3) {
        const F16Image f = f16_image(nullptr, n, a.KH * a.KW, a.Cin, (a.Cout + 15) / 16 * 16);
        return n * 4 * 2 + f.n16 * 2 * 2 + 128;
    }
    return n * 4 * (math_mode == DL4DS_MATH_TF32X3 ? 2 : 1);
}

int conv2d_pack_tc(const float* w, int wmode, int KH, int KW, int Cin, int Cout, int math_mode, void* ws,
                   cudaStream_t st) {
    const Chunk c = pick_chunk(Cin);
    const int nchunks = (Cin + c.kc - 1) / c.kc;
    const int npad = (Cout + 15) / 16 * 16;
    const int64_t n = pack_floats(KH * KW, Cin, Cout);
    float* hi = reinterpret_cast<float*>(ws);
    float* lo = hi + n;
    const int64_t units = n / 4;
    const int blocks = (int)((units + 255) / 256 > 1184 ? 1184 : (units + 255) / 256);
    pack_weights_kernel<<<blocks, 256, 0, st>>>(w, hi, lo, KH * KW, Cin, Cout, npad, c.kc, nchunks, wmode,
                                                math_mode != DL4DS_MATH_TF32 ? 1 : 0);
    if (math_mode == DL4DS_MATH_F16X3) {
        const F16Image f = f16_image(hi, n, KH * KW, Cin, npad);
        weight_absmax_kernel<<<1, 256, 0, st>>>(w, (int64_t)KH * KW * Cin * Cout, f.scale);
        const int64_t u16 = f.n16 / 8;
        const int b16 = (int)((u16 + 255) / 256 > 1184 ? 1184 : (u16 + 255) / 256);
        pack_weights_f16_kernel<<<b16, 256, 0, st>>>(w, hi, n, KH * KW, Cin, Cout, npad, wmode);
    }
    return check_launch("pack_weights_kernel");
}

static std::atomic<bool> g_pack_has_f16{false};

// fills one 64-byte PackDesc (host memory) for conv2d_pack_multi; returns its number of 16-byte units
int64_t conv2d_pack_desc(const float* w, int wmode, int KH, int KW, int Cin, int Cout, int math_mode, void* ws, void* desc_out) {
    const Chunk c = pick_chunk(Cin);
    PackDesc d;
    d.w = w;
    d.taps = KH * KW; d.Cin = Cin; d.Cout = Cout;
    d.Npad = (Cout + 15) / 16 * 16;
    d.kc = c.kc;
    d.nchunks = (Cin + c.kc - 1) / c.kc;
    d.wmode = wmode;
    d.x3 = math_mode == DL4DS_MATH_F16X3 ? 2 : (math_mode == DL4DS_MATH_TF32X3 ? 1 : 0);
    const int64_t n = pack_floats(d.taps, Cin, Cout);
    d.hi = reinterpret_cast<float*>(ws);
    d.lo = d.hi + n;
    d.unit_begin = 0;
    memcpy(desc_out, &d, sizeof(d));
    if (d.x3 == 2) {
        g_pack_has_f16.store(true, std::memory_order_relaxed);
        return n / 4 + f16_image(d.hi, n, d.taps, Cin, d.Npad).n16 / 8;
    }
    return n / 4;
}

int conv2d_pack_multi(const void* descs_dev, int n, int64_t total_units, cudaStream_t st) {
    if (n <= 0 || total_units <= 0) return DL4DS_OK;
    int64_t blocks = (total_units + 255) / 256;
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
    // the weight scales of the fp16 images (DL4DS_MATH_F16X3) first; skipped while no such descriptor was ever built
    if (g_pack_has_f16.load(std::memory_order_relaxed))
        weight_absmax_multi_kernel<<<n, 256, 0, st>>>(reinterpret_cast<const PackDesc*>(descs_dev));
    pack_weights_multi_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const PackDesc*>(descs_dev), n, total_units);
    return check_launch("pack_weights_multi_kernel");
}

int conv2d_fwd_tc(const ConvArgs& a, int math_mode_in, void* ws, int prepacked, cudaStream_t st) {
    // DL4DS_MATH_F16X3: the halo-tile kernel runs on fp16 operands; every other kernel of this file as TF32X3
    const bool f16 = math_mode_in == DL4DS_MATH_F16X3;
    const int math_mode = f16 ? DL4DS_MATH_TF32X3 : math_mode_in;
    if (!fwd_supported(a, math_mode)) return DL4DS_E_UNSUPPORTED;
    DL4DS_REQUIRE(ws != nullptr, DL4DS_E_BADARG,
                  "conv2d_fwd: tensor-core math needs the packed-weight workspace "
                  "(dl4ds_conv2d_fwd_workspace_bytes)");
    DL4DS_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 127) == 0, DL4DS_E_BADARG, "conv2d_fwd: ws must be 128-byte aligned");
    const bool x3 = math_mode == DL4DS_MATH_TF32X3;
    if (!prepacked) {
        int rc = conv2d_pack_tc(a.w, a.wmode, a.KH, a.KW, a.Cin, a.Cout, math_mode_in, ws, st);
        if (rc) return rc;
    }
    const Chunk c = pick_chunk(a.Cin);
    TcFwdParams p;
    p.Npad = (a.Cout + 15) / 16 * 16;
    p.nchunks = (a.Cin + c.kc - 1) / c.kc;
    const int64_t n = pack_floats(a.KH * a.KW, a.Cin, a.Cout);
    p.wp_hi = reinterpret_cast<const float*>(ws);
    p.wp_lo = p.wp_hi + n;
    {   // halo-tile kernel (conv_tc_halo.cu) first; the per-tap kernel below keeps the shapes outside its domain
        if (f16) {
            const F16Image fi = f16_image(reinterpret_cast<float*>(ws), n, a.KH * a.KW, a.Cin, p.Npad);
            const int rc16 = conv2d_fwd_tc_halo(a, DL4DS_MATH_F16X3, reinterpret_cast<const float*>(fi.hi),
                                                reinterpret_cast<const float*>(fi.lo), fi.scale, st);
            if (rc16 != DL4DS_E_UNSUPPORTED) return rc16;
        }
        const int rc = conv2d_fwd_tc_halo(a, math_mode, p.wp_hi, p.wp_lo, nullptr, st);
        if (rc != DL4DS_E_UNSUPPORTED) return rc;
        int bw_, bh_;
        if (!tile_geometry(a.H, a.W, 128, &bw_, &bh_)) return DL4DS_E_UNSUPPORTED;
        if (a.mask_y != nullptr || a.dbias != nullptr) return DL4DS_E_UNSUPPORTED;      // fused epilogue: halo kernel only
    }
    p.bias = a.bias; p.res = a.res; p.y = a.y; p.res_ld = a.res_ld; p.y_ld = a.y_ld;
    p.H = a.H; p.W = a.W; p.Cin = a.Cin; p.Cout = a.Cout;
    p.KW = a.KW; p.ntaps = a.KH * a.KW; p.pad_t = a.pad_t; p.pad_l = a.pad_l;
    tile_geometry(a.H, a.W, 128, &p.BW, &p.BH);
    p.tiles_x = a.W / p.BW;
    p.tiles_per_img = p.tiles_x * (a.H / p.BH);
    p.kc = c.kc; p.span = c.span; p.layout = c.layout;
    p.act = a.act; p.d2s_r = a.d2s_r; p.beta = a.beta;
    p.a_bytes = 128 * c.span;
    p.b_bytes = p.Npad * c.span;
    const int nit = p.ntaps * p.nchunks;
    // x3: the split activations live in TMEM when 2 accumulators + an A ring fit in 512 columns; that
    // removes the A_lo smem copy and all A-operand smem reads of the 3 MMAs (the kernel is otherwise
    // shared-memory-bandwidth bound for narrow N)
    // MEASURED (round 1): correct but slower than the smem path on B200 (SPC dgrad 0.80 vs 0.45 ms), so it is
    // opt-in (DL4DS_TC_A_TMEM=1) until the TMEM-store latency is understood.
    static const bool want_a_tmem = [] { const char* e = getenv("DL4DS_TC_A_TMEM"); return e && e[0] == '1'; }();
    p.a_tmem = (want_a_tmem && x3 && 2 * p.Npad + 2 * 2 * c.kc <= 512) ? 1 : 0;
    static const bool no_stack = [] { const char* e = getenv("DL4DS_TC_NO_STACKN"); return e && e[0] == '1'; }();
    // MEASURED (round 1f): correct (131/131 parity tests) but slower -- 3.47 vs 2.77 ms/step, SPC dgrad 287 vs 116 us: the
    // four row-per-thread splitter warps run one LDS -> split -> tcgen05.st -> wait::st chain per stage, which is longer
    // than the stage's MMAs.  Opt-in (DL4DS_TC_LO_TMEM=1) until the stores are spread over two warps per lane quadrant.
    static const bool want_lo_tmem = [] { const char* e = getenv("DL4DS_TC_LO_TMEM"); return e && e[0] == '1'; }();
    p.stackn = (x3 && !p.a_tmem && p.Npad <= 64 && !no_stack) ? 1 : 0;
    p.lo_tmem = (p.stackn && want_lo_tmem) ? 1 : 0;
    const int it_bytes = (p.a_tmem || p.lo_tmem) ? (p.a_bytes + 2 * p.b_bytes) : (p.a_bytes + p.b_bytes) * (x3 ? 2 : 1);
    int group = (32 * 1024) / it_bytes;                    // ~32 KB per stage: few barrier round trips per tile
    if (group < 1) group = 1;
    if (group > nit) group = nit;
    if (p.a_tmem && group * c.kc > 32) group = 32 / c.kc;  // the splitter holds one stage row (<= 32 floats) in registers
    if (p.lo_tmem) group = 1;                              // one chunk per stage: TMEM columns (kc per stage) decide the ring depth
    p.group = group;
    p.ngroups = (nit + group - 1) / group;
    p.stage_bytes = group * it_bytes;
    p.acc_stride = p.stackn ? 2 * p.Npad : p.Npad;
    p.tmem_a_off = 2 * p.acc_stride;                       // the A ring follows the two accumulator buffers
    int acc_cols = 32;
    while (acc_cols < 2 * p.acc_stride) acc_cols *= 2;     // double-buffered accumulator alone
    // persistent CTAs: two per SM when TMEM (<= 256 columns each) allows, else one with all the smem
    int ctas_per_sm = acc_cols <= 256 ? 2 : 1;
    int stages;
    int cols;
    for (;;) {
        stages = ((ctas_per_sm == 2 ? 104 : 208) * 1024) / p.stage_bytes;
        if (stages > kMaxStages) stages = kMaxStages;
        const int budget = (ctas_per_sm == 2 ? 256 : 512) - 2 * p.acc_stride;
        if (p.a_tmem || p.lo_tmem) {
            const int by_tmem = budget / (group * (p.a_tmem ? 2 : 1) * c.kc);
            if (stages > by_tmem) stages = by_tmem;
        }
        if (stages >= (p.lo_tmem ? 3 : 2) || ctas_per_sm == 1) break;   // measured: 2 CTAs x 2 stages beat 1 CTA x 4 stages
        ctas_per_sm = 1;
    }
    if (stages < 2) stages = 2;
    p.stages = stages;
    cols = 32;
    while (cols < 2 * p.acc_stride + (p.a_tmem ? stages * group * 2 * c.kc : (p.lo_tmem ? stages * group * c.kc : 0))) cols *= 2;
    if (cols > (ctas_per_sm == 2 ? 256 : 512)) { p.a_tmem = 0; return DL4DS_E_UNSUPPORTED; }
    p.tmem_cols = cols;
    const size_t smem = (size_t)stages * p.stage_bytes + 1024;
    DL4DS_REQUIRE(smem <= 220 * 1024, DL4DS_E_UNSUPPORTED, "conv2d_fwd_tc: stage too large");
    const CUtensorMap* tm = get_tensor_map_nhwc(a.x, a.x_ld, a.N, a.H, a.W, a.Cin, c.kc, p.BW, p.BH, c.swz);
    if (!tm) return DL4DS_E_CUDA;
    p.ntiles = a.N * p.tiles_per_img;
    const int grid = p.ntiles < ctas_per_sm * kNumSMs ? p.ntiles : ctas_per_sm * kNumSMs;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(conv_tc_fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(221 * 1024));
        cudaFuncSetAttribute(conv_tc_fwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(221 * 1024));
        cudaFuncSetAttribute(conv_tc_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(221 * 1024));
        attr_done = true;
    }
    if (x3 && (p.a_tmem || p.lo_tmem))
        conv_tc_fwd_kernel<true, true><<<grid, kTcThreadsX3, smem, st>>>(*tm, p);
    else if (x3)
        conv_tc_fwd_kernel<true, false><<<grid, kTcThreadsX3, smem, st>>>(*tm, p);
    else
        conv_tc_fwd_kernel<false, false><<<grid, kTcThreads, smem, st>>>(*tm, p);
    g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    return check_launch("conv_tc_fwd_kernel");
}

// -------------------------------------------------------------------------------------------------
// weight-gradient kernel:  dw[kh][kw][ca][cb] += sum_{n,y,x} P[n, y+kh-pad_t, x+kw-pad_l, ca] * Q[n,y,x,cb]
//
// Per kernel tap a GEMM whose reduction dimension is the PIXELS: D (Ca x Cb) += P_tap^T (Ca x pix) *
// Q (pix x Cb).  kind::tf32 only accepts K-major operands (measured: the MN-major bits of the
// instruction descriptor yield zeros for tf32), i.e. rows = channels, 32 consecutive pixels per
// 128-byte row -- the transpose of what NHWC delivers.  So each stage is:
//   TMA: NHWC boxes {KC ch, BW, BH, 1} of 32 pixels (P shifted by the tap, Q unshifted) -> raw region
//   4 transposer warps: swizzle-aware conflict-free LDS.128 of (pixel, 4 channels), tf32 hi/lo split
//     in registers (x3 mode), STS.32 into the channel-major SWIZZLE_128B operand tiles
//   1 thread: tcgen05.mma M=64 (<= 64 input channels; rows past Ca hold stale data whose D rows are
//     never read), N = block of <= 128 output channels, 4 K-steps of 8 pixels per stage and tap.
// CTA role (blockIdx.y) = (kernel row kh, Ca group, Cb block): it owns the KW accumulators of that
// kernel row in TMEM (KW x Nmma columns) so one Q tile feeds KW taps.  blockIdx.x splits the pixel
// tiles (split-K); partial sums are merged with fp32 atomics.
// -------------------------------------------------------------------------------------------------
constexpr int kWgThreads = 320;       // warp 0 TMA, warp 1 MMA, warps 2-9 transposers (2-5 also epilogue)
constexpr int kWgMaxUnits = 128;      // (raw box, 4-channel group) pairs per stage

struct TcWgradParams {
    float* dw;
    int H, W, Ca, Cb;
    int KH, KW, pad_t, pad_l;
    int nk;                           // 32-pixel chunks per stage (KT = 32*nk pixels)
    int BW, BH, tiles_x, tiles_per_img, ntiles, tiles_per_split;
    int kc_p, span_p, kc_q, span_q;
    int ncig, ncob, Nb;
    int nblk_p_max, nblk_q_max;       // raw boxes per tap (P) / per tile (Q) the smem layout is sized for
    int box_p, box_q;                 // bytes per raw TMA box (KT pixels x span)
    int raw_bytes;                    // raw region per stage
    int pt_bytes, qt_bytes;           // transposed tiles per 32-pixel chunk: P^T per tap, Q^T
    int chunk_bytes;                  // KW*pt_bytes + qt_bytes: operand tiles of one 32-pixel chunk
    int op_bytes;                     // nk*chunk_bytes: one (hi) operand set
    int rstages, ostages;             // raw (TMA) ring depth, operand-tile ring depth
    int ops_set_bytes;                // op_bytes * (1 or 2): operand bytes of one ring slot
    int ops_base;                     // byte offset of the operand ring (after the raw ring)
    int tmem_cols;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <bool X3>
__global__ void __launch_bounds__(kWgThreads) conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_p,
                                                                  const __grid_constant__ CUtensorMap tmap_q,
                                                                  const TcWgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    // raw ring: TMA -> transposers (full) and back (rfree); operand ring: transposers -> MMA (conv) and back (empty)
    __shared__ __align__(8) uint64_t bar_full[kMaxStages];
    __shared__ __align__(8) uint64_t bar_rfree[kMaxStages];
    __shared__ __align__(8) uint64_t bar_conv[kMaxStages];
    __shared__ __align__(8) uint64_t bar_empty[kMaxStages];
    __shared__ __align__(8) uint64_t bar_accum;
    __shared__ uint32_t tmem_base_smem;
    __shared__ uint4 unit_tab[kWgMaxUnits];     // {raw box offset, group, dst tile offset in a chunk, first row | isQ<<16}

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const smem_al = smem_raw + (smem_base - smem_u32(smem_raw));

    // role
    const int role = blockIdx.y;
    const int cob = role % p.ncob;
    const int cig = (role / p.ncob) % p.ncig;
    const int kh = role / (p.ncob * p.ncig);
    const int ca0 = cig * 64;
    const int ca_n = min(64, p.Ca - ca0);
    const int nblk_p = ca_n / p.kc_p;
    const int cb0 = cob * p.Nb;
    const int cb_n = min(p.Nb, p.Cb - cb0);
    const int nblk_q = cb_n / p.kc_q;
    const int Nmma = (cb_n + 15) & ~15;
    // pixel-tile range of this split
    const int t_begin = blockIdx.x * p.tiles_per_split;
    const int t_end = min(p.ntiles, t_begin + p.tiles_per_split);
    const int my_tiles = t_end - t_begin;
    if (my_tiles <= 0) return;

    const uint32_t rawq_off = (uint32_t)(p.KW * p.nblk_p_max * p.box_p);     // Q boxes inside the raw region
    const uint32_t qt_off = (uint32_t)(p.KW * p.pt_bytes);                  // Q^T inside a chunk's operand tiles
    const int gp = p.kc_p / 4, gq = p.kc_q / 4;
    const int units_p = p.KW * nblk_p * gp;
    const int nunits = units_p + nblk_q * gq;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.rstages; ++s) {
            mbar_init(smem_u32(&bar_full[s]), 1);
            mbar_init(smem_u32(&bar_rfree[s]), 8);
        }
        for (int s = 0; s < p.ostages; ++s) {
            mbar_init(smem_u32(&bar_conv[s]), 8);
            mbar_init(smem_u32(&bar_empty[s]), 1);
        }
        mbar_init(smem_u32(&bar_accum), 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmap_p);
        tma_prefetch_desc(&tmap_q);
    }
    for (int u = threadIdx.x; u < nunits; u += blockDim.x) {
        uint4 e;
        if (u < units_p) {
            const int g = u % gp, tb = u / gp;
            const int kw = tb / nblk_p, b = tb - kw * nblk_p;
            e.x = (uint32_t)((kw * p.nblk_p_max + b) * p.box_p);
            e.y = (uint32_t)g;
            e.z = (uint32_t)(kw * p.pt_bytes);
            e.w = (uint32_t)(b * p.kc_p + g * 4);
        } else {
            const int j = u - units_p;
            const int g = j % gq, b = j / gq;
            e.x = rawq_off + (uint32_t)(b * p.box_q);
            e.y = (uint32_t)g;
            e.z = qt_off;
            e.w = (uint32_t)(b * p.kc_q + g * 4) | (1u << 16);
        }
        unit_tab[u] = e;
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_smem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const uint32_t tx_bytes = (uint32_t)(p.KW * nblk_p * p.box_p + nblk_q * p.box_q);
            for (int it = 0; it < my_tiles; ++it) {
                const int s = it % p.rstages;
                const uint32_t ph = (uint32_t)((it / p.rstages) & 1);
                mbar_wait(smem_u32(&bar_rfree[s]), ph ^ 1u);
                const uint32_t full = smem_u32(&bar_full[s]);
                mbar_arrive_expect_tx(full, tx_bytes);
                const int tile = t_begin + it;
                const int img = tile / p.tiles_per_img;
                const int trem = tile - img * p.tiles_per_img;
                const int ty = trem / p.tiles_x, tx = trem - ty * p.tiles_x;
                const int y0 = ty * p.BH, x0 = tx * p.BW;
                const uint32_t sp = smem_base + (uint32_t)s * (uint32_t)p.raw_bytes;
                for (int kw = 0; kw < p.KW; ++kw)
                    for (int b = 0; b < nblk_p; ++b)
                        tma_load_4d(sp + (uint32_t)((kw * p.nblk_p_max + b) * p.box_p), &tmap_p, full,
                                    ca0 + b * p.kc_p, x0 + kw - p.pad_l, y0 + kh - p.pad_t, img);
                for (int b = 0; b < nblk_q; ++b)
                    tma_load_4d(sp + rawq_off + (uint32_t)(b * p.box_q), &tmap_q, full, cb0 + b * p.kc_q, x0, y0, img);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(64, Nmma, 0, 0);
            for (int it = 0; it < my_tiles; ++it) {
                const int s = it % p.ostages;
                const uint32_t ph = (uint32_t)((it / p.ostages) & 1);
                mbar_wait(smem_u32(&bar_conv[s]), ph);
                tc_fence_after();
                const uint32_t so = smem_base + (uint32_t)p.ops_base + (uint32_t)s * (uint32_t)p.ops_set_bytes;
                for (int j = 0; j < p.nk; ++j) {
                    const uint32_t sc = so + (uint32_t)(j * p.chunk_bytes);
                    const uint32_t qb = sc + qt_off;
                    for (int kw = 0; kw < p.KW; ++kw) {
                        const uint32_t pa = sc + (uint32_t)(kw * p.pt_bytes);
                        const uint32_t td = tmem_d + (uint32_t)(kw * Nmma);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t acc = (it > 0 || j > 0 || k > 0) ? 1u : 0u;
                            const uint32_t ko = (uint32_t)k * 32u;
                            const uint64_t da = make_smem_desc(pa + ko, 16, 1024, kLayoutSw128);
                            const uint64_t db = make_smem_desc(qb + ko, 16, 1024, kLayoutSw128);
                            if (X3) {
                                const uint32_t h = (uint32_t)p.op_bytes;
                                const uint64_t dal = make_smem_desc(pa + h + ko, 16, 1024, kLayoutSw128);
                                const uint64_t dbl = make_smem_desc(qb + h + ko, 16, 1024, kLayoutSw128);
                                umma_tf32(td, dal, db, idesc, acc);
                                umma_tf32(td, da, dbl, idesc, 1u);
                                umma_tf32(td, da, db, idesc, 1u);
                            } else {
                                umma_tf32(td, da, db, idesc, acc);
                            }
                        }
                    }
                }
                umma_commit(smem_u32(&bar_empty[s]));
            }
            umma_commit(smem_u32(&bar_accum));
        }
    } else {
        // ===================== transposers (+ tf32 split), then epilogue on warps 2-5 =====================
        const int tw = warp - 2;                       // 0..7
        // lane-dependent parts of the swizzled addresses (this thread always handles pixel row `lane` of a chunk)
        const uint32_t src_lane_p = (uint32_t)(lane * p.span_p), src_lane_q = (uint32_t)(lane * p.span_q);
        const uint32_t dst_lane = (uint32_t)((lane & 3) << 2);
        const uint32_t lane_unit = (uint32_t)(lane >> 2);
        for (int it = 0; it < my_tiles; ++it) {
            const int rs = it % p.rstages, os = it % p.ostages;
            mbar_wait(smem_u32(&bar_full[rs]), (uint32_t)((it / p.rstages) & 1));
            mbar_wait(smem_u32(&bar_empty[os]), (uint32_t)(((it / p.ostages) & 1) ^ 1));
            uint8_t* raw = smem_al + (size_t)rs * p.raw_bytes;
            uint8_t* ops = smem_al + p.ops_base + (size_t)os * p.ops_set_bytes;
            for (int j = 0; j < p.nk; ++j)
            for (int u = tw; u < nunits; u += 8) {
                const uint4 e = unit_tab[u];
                const bool isq = (e.w >> 16) != 0;
                const int span = isq ? p.span_q : p.span_p;
                const uint32_t chunk_rows = (uint32_t)(j * 32) * (uint32_t)span;
                const uint8_t* src = raw + e.x + chunk_rows + (isq ? src_lane_q : src_lane_p) +
                                     ((uint32_t)swizzle_unit((int)e.y, lane, span) << 4);
                const float4 v = *reinterpret_cast<const float4*>(src);
                uint8_t* dst = ops + (size_t)j * p.chunk_bytes + e.z;
                const uint32_t row0 = e.w & 0xffffu;
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const uint32_t row = row0 + r;
                    const uint32_t off = row * 128u + (((lane_unit ^ (row & 7u)) << 4) | dst_lane);
                    if (X3) {
                        const float h = tf32_rna(vv[r]);
                        *reinterpret_cast<float*>(dst + off) = h;
                        *reinterpret_cast<float*>(dst + p.op_bytes + off) = tf32_rna(vv[r] - h);
                    } else {
                        *reinterpret_cast<float*>(dst + off) = vv[r];
                    }
                }
            }
            mbar_arrive_warp(smem_u32(&bar_rfree[rs]));      // raw slot read: the next TMA may overwrite it
            fence_proxy_async_smem();
            mbar_arrive_warp(smem_u32(&bar_conv[os]));
        }
        if (tw < 4) {
            mbar_wait(smem_u32(&bar_accum), 0);
            tc_fence_after();
            // M = 64 accumulator layout: row r lives in TMEM lane 32*(r/16) + r%16
            const int q = warp & 3;
            const int ci = ca0 + q * 16 + lane;                 // valid for lane < 16
            const bool row_ok = lane < 16 && (q * 16 + lane) < ca_n;
            const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16);
            for (int kw = 0; kw < p.KW; ++kw) {
                float* dst_row = p.dw + ((int64_t)((kh * p.KW + kw) * p.Ca + ci)) * p.Cb + cb0;
                for (int c0 = 0; c0 < Nmma; c0 += 16) {
                    float v[16];
                    tmem_ld16(taddr + (uint32_t)(kw * Nmma + c0), v);
                    if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            if (c0 + j < cb_n) red_add_v4(dst_row + c0 + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, (uint32_t)p.tmem_cols);
}

static bool wgrad_supported(const WgradArgs& a, int math_mode) {
    if (math_mode != DL4DS_MATH_TF32 && math_mode != DL4DS_MATH_TF32X3) return false;
    if (!is_sm100()) return false;
    if (a.stride != 1 || a.Hp != a.Hq || a.Wp != a.Wq) return false;
    if (a.Ca % 8 || a.Cb % 8) return false;
    if (a.p_ld % 4 || a.q_ld % 4) return false;
    if ((reinterpret_cast<uintptr_t>(a.P) & 15) || (reinterpret_cast<uintptr_t>(a.Q) & 15)) return false;
    if (a.KW > 5 || a.KH > 9) return false;
    if ((reinterpret_cast<uintptr_t>(a.dw) & 15)) return false;
    int bw, bh;
    return tile_geometry(a.Hq, a.Wq, 32, &bw, &bh);
}

int64_t conv2d_wgrad_tc_workspace(int, int, int, int, int, int, int) { return 0; }

int conv2d_wgrad_tc(const WgradArgs& a, void*, int math_mode, cudaStream_t st) {
    if (!wgrad_supported(a, math_mode)) return DL4DS_E_UNSUPPORTED;
    const bool x3 = math_mode == DL4DS_MATH_TF32X3;
    const Chunk cp = pick_chunk(a.Ca), cq = pick_chunk(a.Cb);
    TcWgradParams p;
    // pixels per stage: 32*nk, aiming at >= 16 KB of raw TMA traffic per stage (latency amortisation)
    const int ca_f = a.Ca < 64 ? a.Ca : 64;
    const int raw32 = 32 * 4 * (a.KW * ca_f + (a.Cb < 128 ? a.Cb : 128));
    int nk = (16 * 1024 + raw32 - 1) / raw32;
    if (nk > 8) nk = 8;
    int BW = 0, BH = 0;
    while (nk > 1 && !tile_geometry(a.Hq, a.Wq, 32 * nk, &BW, &BH)) --nk;
    tile_geometry(a.Hq, a.Wq, 32 * nk, &BW, &BH);
    p.nk = nk;
    p.dw = a.dw;
    p.H = a.Hq; p.W = a.Wq; p.Ca = a.Ca; p.Cb = a.Cb;
    p.KH = a.KH; p.KW = a.KW; p.pad_t = a.pad_t; p.pad_l = a.pad_l;
    p.BW = BW; p.BH = BH;
    p.tiles_x = a.Wq / BW;
    p.tiles_per_img = p.tiles_x * (a.Hq / BH);
    p.ntiles = a.N * p.tiles_per_img;
    p.kc_p = cp.kc; p.span_p = cp.span;
    p.kc_q = cq.kc; p.span_q = cq.span;
    p.ncig = (a.Ca + 63) / 64;
    // output-channel block: <= 128 (<= 96 for KW = 5 so that KW*Nmma <= 512 TMEM columns), a multiple of
    // lcm(16, kc_q), balanced over the blocks
    const int nb_cap = a.KW > 3 ? 96 : 128;
    const int unit = cq.kc > 16 ? cq.kc : 16;
    int ncob = (a.Cb + nb_cap - 1) / nb_cap;
    int nb = ((a.Cb + ncob - 1) / ncob + unit - 1) / unit * unit;
    if (nb > nb_cap) { nb = nb_cap / unit * unit; ncob = (a.Cb + nb - 1) / nb; }
    p.Nb = nb; p.ncob = ncob;
    const int ca_first = a.Ca < 64 ? a.Ca : 64;
    p.nblk_p_max = ca_first / cp.kc;
    const int cb_first = a.Cb < nb ? a.Cb : nb;
    p.nblk_q_max = cb_first / cq.kc;
    p.box_p = 32 * nk * cp.span;
    p.box_q = 32 * nk * cq.span;
    p.raw_bytes = a.KW * p.nblk_p_max * p.box_p + p.nblk_q_max * p.box_q;       // multiples of 1024
    const int nmma_max = (cb_first + 15) & ~15;
    p.pt_bytes = ((ca_first + 7) & ~7) * 128;              // rows past Ca alias whatever follows (never read back)
    p.qt_bytes = nmma_max * 128;
    p.chunk_bytes = a.KW * p.pt_bytes + p.qt_bytes;
    p.op_bytes = nk * p.chunk_bytes;
    p.ops_set_bytes = p.op_bytes * (x3 ? 2 : 1);
    if (a.KW * (ca_first / cp.kc) * (cp.kc / 4) + (cb_first / cq.kc) * (cq.kc / 4) > kWgMaxUnits) return DL4DS_E_UNSUPPORTED;
    const int slack = 64 * 128;                            // the M=64 operand window of the last tile stays in-bounds
    int cols = 32;
    while (cols < a.KW * nmma_max) cols *= 2;
    if (cols > 512) return DL4DS_E_UNSUPPORTED;
    p.tmem_cols = cols;
    const int nroles = a.KH * p.ncig * p.ncob;
    // ring depths: operand ring 2 deep (transpose of tile i+1 overlaps the MMAs of tile i), the rest of the
    // budget goes to the raw TMA ring (up to 4) to cover the global-load latency
    int budget = 218 * 1024 - slack;
    int splits = kNumSMs / nroles;
    if (cols <= 256 && 2 * p.ops_set_bytes + 3 * p.raw_bytes <= 100 * 1024) {   // two CTAs per SM
        budget = 100 * 1024;
        splits = (2 * kNumSMs) / nroles;
    }
    int ostages = 2;
    if (budget < ostages * p.ops_set_bytes + p.raw_bytes) ostages = 1;
    if (budget < ostages * p.ops_set_bytes + p.raw_bytes) return DL4DS_E_UNSUPPORTED;
    int rstages = (budget - ostages * p.ops_set_bytes) / p.raw_bytes;
    if (rstages > 4) rstages = 4;
    if (splits < 1) splits = 1;
    if (splits > p.ntiles) splits = p.ntiles;
    p.tiles_per_split = (p.ntiles + splits - 1) / splits;
    splits = (p.ntiles + p.tiles_per_split - 1) / p.tiles_per_split;
    p.rstages = rstages;
    p.ostages = ostages;
    p.ops_base = rstages * p.raw_bytes;
    const size_t smem = (size_t)p.ops_base + (size_t)ostages * p.ops_set_bytes + slack + 1024;
    const CUtensorMap* tp = get_tensor_map_nhwc(a.P, a.p_ld, a.N, a.Hp, a.Wp, a.Ca, cp.kc, BW, BH, cp.swz);
    const CUtensorMap* tq = get_tensor_map_nhwc(a.Q, a.q_ld, a.N, a.Hq, a.Wq, a.Cb, cq.kc, BW, BH, cq.swz);
    if (!tp || !tq) return DL4DS_E_CUDA;
    dim3 grid((unsigned)splits, (unsigned)nroles);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(conv_tc_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(221 * 1024));
        cudaFuncSetAttribute(conv_tc_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(221 * 1024));
        attr_done = true;
    }
    if (x3)
        conv_tc_wgrad_kernel<true><<<grid, kWgThreads, smem, st>>>(*tp, *tq, p);
    else
        conv_tc_wgrad_kernel<false><<<grid, kWgThreads, smem, st>>>(*tp, *tq, p);
    g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    return check_launch("conv_tc_wgrad_kernel");
}

}  // namespace dl4ds
