// sm_100a building blocks for the tensor-core convolution kernels (conv_tc.cu): mbarrier, TMA
// (cp.async.bulk[.tensor]), tcgen05 (alloc / mma kind::tf32 / commit / ld) wrappers as inline PTX,
// UMMA shared-memory and instruction descriptors, and the host-side CUtensorMap cache.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace dl4ds {
namespace tc {

// UMMA shared-memory layout types (descriptor bits [61,64))
constexpr uint32_t kLayoutSw128 = 2, kLayoutSw64 = 4, kLayoutSw32 = 6;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// one arrival per WARP: every lane's prior shared-memory accesses are ordered before lane 0's
// (release) arrive by the __syncwarp.  512 per-thread arrivals on one mbarrier serialise in the
// shared-memory atomic unit (measured ~1-2k cycles per phase); barriers fed by whole warps are
// initialised with the warp count and use this.
__device__ __forceinline__ void mbar_arrive_warp(uint32_t bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
// plain bulk copy global -> shared (size multiple of 16, both 16-byte aligned)
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
        : "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], kind::tf32 (fp32 storage, 10-bit mantissa operands, fp32 accumulate)
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same with the A operand in TMEM (lane = row m, one 32-bit column per k element)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 8 consecutive 32-bit columns, registers -> TMEM (thread i writes lane base_lane + i)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                   "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp gets lane (base_lane + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, version 1 = sm_100)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    d |= static_cast<uint64_t>(layout_type) << 61;
    return d;
}
// instruction descriptor for kind::tf32, fp32 accumulate (cute::UMMA::InstrDescriptor bit layout)
__host__ __device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                       // c_format = F32
    d |= 2u << 7;                       // a_format = TF32
    d |= 2u << 10;                      // b_format = TF32
    d |= static_cast<uint32_t>(a_mn_major & 1) << 15;
    d |= static_cast<uint32_t>(b_mn_major & 1) << 16;
    d |= static_cast<uint32_t>(N >> 3) << 17;
    d |= static_cast<uint32_t>(M >> 4) << 24;
    return d;
}

// instruction descriptor for kind::f16 with fp16 operands, fp32 accumulate (a_format = b_format = 0 = F16)
__host__ __device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                       // c_format = F32
    d |= static_cast<uint32_t>(N >> 3) << 17;
    d |= static_cast<uint32_t>(M >> 4) << 24;
    return d;
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16 (K = 16 per instruction: 32-byte K-major rows)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// round-to-nearest (ties away from zero in magnitude) onto the tf32 grid with two integer ops: add half
// an ulp of the 10-bit mantissa, clear the 13 low bits.  (cvt.rna.tf32.f32 expands to ~8 SASS
// instructions with Inf/NaN handling; the operand splitters run this per element.)  Inf/NaN inputs
// are not expected on this path.
__device__ __forceinline__ float tf32_rna(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
// what the tensor core itself does with an fp32 operand of kind::tf32: keep sign, exponent and 10 mantissa bits
__device__ __forceinline__ float tf32_trunc(float x) {
    return __uint_as_float(__float_as_uint(x) & 0xffffe000u);
}
// lo operand of the raw-tile-as-hi scheme: v - trunc(v) is exact but carries up to 13 significant bits, of which the
// tensor core would TRUNCATE the last two (a one-signed error of up to 2^-21 |v| per element that accumulates coherently
// over a reduction); rounding it onto the tf32 grid here makes the residual error half as large and unbiased.
__device__ __forceinline__ float tf32_lo_of_trunc(float v) {
    return tf32_rna(v - tf32_trunc(v));
}
// hi/lo split for the 3xTF32 scheme: x = hi + lo exactly; the tensor core truncates lo to tf32
// (|lo| <= 2^-12 |x|, so the truncation error is <= 2^-22 |x|).
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) {
    hi = tf32_rna(x);
    lo = x - hi;
}

// channel-chunk geometry shared by the pack kernel, the tensor maps and the UMMA descriptors:
// KC channels (8 / 16 / 32 fp32 = one 32 / 64 / 128-byte swizzle span) per pixel row.
struct Chunk {
    int kc;          // channels per chunk
    int span;        // bytes per row = kc*4
    uint32_t layout; // UMMA layout type
    int swz;         // CUtensorMapSwizzle value
};
static inline Chunk pick_chunk(int C) {
    if (C % 32 == 0) return {32, 128, kLayoutSw128, (int)CU_TENSOR_MAP_SWIZZLE_128B};
    if (C % 16 == 0) return {16, 64, kLayoutSw64, (int)CU_TENSOR_MAP_SWIZZLE_64B};
    return {8, 32, kLayoutSw32, (int)CU_TENSOR_MAP_SWIZZLE_32B};
}
// 2^(13 - floor(log2(amax))): amax * s lands in [2^13, 2^14); 1 for zero / denormal / non-finite maxima.  The scale of
// the 3-term fp16 mode: a power of two, so scaling and un-scaling are exact.
__device__ __forceinline__ float pow2_scale_for(float amax) {
    const int e = (int)((__float_as_uint(amax) >> 23) & 0xffu);
    if (e == 0 || e == 255) return 1.0f;
    int se = 127 + 13 - (e - 127);
    se = se < 1 ? 1 : (se > 254 ? 254 : se);
    return __uint_as_float((uint32_t)se << 23);
}

// the same for fp16 operands (3-term fp16 mode): channels are padded to a multiple of 16 (K = 16 per MMA), KC halfs =
// one 32 / 64 / 128-byte span
struct Chunk16 {
    int kc, span, nchunks, cpad;
    uint32_t layout;
};
static inline Chunk16 pick_chunk16(int C) {
    const int cp = (C + 15) / 16 * 16;
    if (cp % 64 == 0) return {64, 128, cp / 64, cp, kLayoutSw128};
    if (cp % 32 == 0) return {32, 64, cp / 32, cp, kLayoutSw64};
    return {16, 32, cp / 16, cp, kLayoutSw32};
}
// 16-byte-unit XOR of the hardware swizzle for row r of a tile whose rows are `span` bytes
__host__ __device__ __forceinline__ int swizzle_unit(int unit, int row, int span) {
    if (span == 128) return unit ^ (row & 7);
    if (span == 64) return unit ^ ((row >> 1) & 3);
    return unit ^ ((row >> 2) & 1);
}

// host: cached 4-D NHWC tensor map {C, W, H, N} with box {kc, bw, bh, 1}
const CUtensorMap* get_tensor_map_nhwc(const float* base, int ld, int N, int H, int W, int C,
                                       int box_c, int box_w, int box_h, int swizzle);

}  // namespace tc
}  // namespace dl4ds
