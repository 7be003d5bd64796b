// tcgen05 weight-gradient kernel, second generation ("stacked taps"):
//
//   dw[kh][kw][ca][cb] += sum_{n,y,x} P[n, y+kh-pad_t, x+kw-pad_l, ca] * Q[n, y, x, cb]
//
// (what TF's GradientTape derives for every stride-1 Conv2D of blocks.py:49-61,208,299,414-416 and
// sp_postups.py:134,156.)  The reduction dimension of this GEMM is the PIXELS, and kind::tf32 only
// takes K-major operands, so NHWC data has to be transposed on chip (pixel-major -> channel-major).
//
// What r01b's profile showed for the first-generation kernel (conv_tc.cu): M=64 instruction shape =
// half tensor rate with 48 of 64 rows used (2.67x the forward's tensor cycles), and one CTA role per
// kernel row, i.e. every pixel re-loaded and re-transposed KH times.  This kernel instead
//   * stacks ALL taps on the M dimension: A rows = [(kh,kw)][ca], KH*KW*Ca rows cut into M=128
//     blocks (432 rows -> 4 blocks for the 48-channel 3x3 layers, 84 % of full-rate rows);
//   * B = Q^T (N = a block of <= 512/nblk output channels), one fp32 accumulator per M block in TMEM;
//   * loads P once per 32-pixel chunk as ONE halo box per channel chunk ({kc, BW+KW-1, BH+KH-1}
//     TMA box, out-of-bounds = the convolution's zero padding) instead of KH*KW shifted boxes;
//   * hands the MMA warp one 128-row block at a time through a ring of operand slots, so the
//     transposition of the next blocks overlaps the MMAs of this one (clock64 stamps of two earlier
//     layouts -- all warps sharing a block with one unit in flight each, and one private slot per
//     warp group -- showed the transposers latency-bound resp. in lock-step with the MMA warp).
//
// Operand orientation.  scratch/umma_rate.cu measured the cost of one kind::tf32 M=128 K=8 MMA with both
// operands in shared memory: ~119 cycles for ANY N <= 128, 139 at N=192, 171 at N=256 -- a per-instruction
// floor, so the instruction count is what matters and N should be as large as TMEM allows:
//   * Cb <= 128 ("swap"): A = Q^T (one M tile, rows past Cb are never read back), B = a block of up to 256
//     stacked P^T rows; 432 stacked rows -> 2 blocks of 224/208 -> 24 MMAs per 32-pixel chunk (x3) instead
//     of 48 with the stacked rows on M;
//   * Cb > 128 (the 48 -> 192 sub-pixel layer): A = 128-row block of stacked P^T, B = Q^T with N = Cb (<= 256);
//     the blocks are spread over `nrg` CTA roles because bpr * N accumulator columns must fit TMEM.
//
// Warp roles (kWg2Threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-5 = Q
// group (transposes the Q^T tile of every chunk), warps 6-17 = A group (every block of the stacked
// P^T, 3 units in flight per warp).  A transposer unit = swizzle-aware LDS.128 of (pixel, 4
// channels), tf32 hi/lo split, STS.32 into the channel-major SWIZZLE_128B operand tile.  All 16
// transposer warps drain TMEM at the end (red.global.add.v4).
// blockIdx.y = (input-channel group, output-channel block) role, blockIdx.x = split of the pixel
// chunks (split-K); partial sums meet in fp32 atomics on the gradient arena.
#include <stdlib.h>

#include "tc_common.cuh"

namespace dl4ds {

using namespace tc;

constexpr int kWg2TransWarps = 16;     // split into a Q group (qwarps) and an A group (the rest) per launch
constexpr int kWg2Threads = (2 + kWg2TransWarps) * 32;
constexpr int kWg2MaxAUnits = 128;     // 512 stacked rows / 4
constexpr int kWg2MaxQUnits = 64;      // 256 output channels / 4
constexpr int kWg2MaxStages = 8;

struct Wg2Params {
    float* dw;
    int H, W, Ca, Cb, KH, KW, pad_t, pad_l;
    int BW, BH, PW;                   // 32-pixel chunk geometry, halo row pitch (BW + KW - 1)
    int tiles_x, tiles_per_img, ntiles, tiles_per_split;
    int kc_p, span_p, kc_q, span_q;
    int CaG, ncig, Nb, ncob;
    int swap;                         // 1: A = Q^T, B = stacked-row block (N = BR); 0: A = stacked-row block, B = Q^T
    int BR;                           // stacked rows per block (128 when !swap, <= 256 when swap)
    int bpr, nrg;                     // blocks per CTA role, row-group roles
    int qwarps, qstages;
    int stackm;                       // swap, x3, 2*q_rows <= 128: A = [Q_hi ; Q_lo] stacked on M -> 2 MMAs per K-step
    int q_rows;                       // rows of one Q^T half in its slot
    int box_p, box_q;                 // smem bytes reserved per raw box (multiples of 1024)
    int tx_p, tx_q;                   // bytes one TMA box transfers
    int nbox_p_max;                   // P boxes of a full input-channel group (Q boxes follow them)
    int raw_bytes;
    int a_half, q_half;               // bytes of the hi half of an A / Q slot (lo follows in x3 mode)
    int a_slot, q_slot;
    int rstages, astages;             // raw (TMA) ring depth, A-slot ring depth
    int q_base, raw_base, a_base;     // smem byte offsets: Q slots first (an A-operand read of a short Q^T tile
                                      // runs on into the raw ring, never out of bounds), raw ring, A-slot ring
    int tmem_cols;
    long long* dbg;                   // optional: clock64 stamps of CTA (0,0) (dl4ds_debug_set_buffer), else NULL
};

// debug stamps: slot = it * 16 + id, only CTA (0,0), only the first 64 chunks
#define WG2_STAMP(id)                                                                           \
    do {                                                                                        \
        if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && it < 64 && lane == 0)     \
            p.dbg[it * 16 + (id)] = clock64();                                                  \
    } while (0)

// whole-kernel stamps of CTA (0,0): slot 1024 + id (0 entry, 1 setup done, 2 main loop drained, 3 epilogue done)
#define WG2_KSTAMP(id)                                                                          \
    do {                                                                                        \
        if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 64)        \
            p.dbg[1024 + (id)] = clock64();                                                     \
    } while (0)

__device__ __forceinline__ void red_add_v4_f32(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// One unit = 32 pixels (lane = pixel of the chunk) x 4 consecutive channels: read the 16 bytes of this
// lane's pixel through the TMA swizzle, write 4 channel-major rows (K-major SWIZZLE_128B, 32 pixels =
// 128 bytes per row).  Both sides are bank-conflict free.
template <bool X3>
__device__ __forceinline__ void transpose_unit(const uint8_t* box, int span, int src_row, int g, uint8_t* tile,
                                               int lo_off, int row0, int lane) {
    const int x = ((src_row * span) >> 7) & ((span >> 4) - 1);
    const float4 v = *reinterpret_cast<const float4*>(box + src_row * span + ((g ^ x) << 4));
    const float vv[4] = {v.x, v.y, v.z, v.w};
    const uint32_t lu = (uint32_t)(lane >> 2), lb = (uint32_t)((lane & 3) << 2);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t row = (uint32_t)(row0 + r);
        const uint32_t off = row * 128u + (((lu ^ (row & 7u)) << 4) | lb);
        if (X3) {
            const float h = tf32_rna(vv[r]);
            *reinterpret_cast<float*>(tile + off) = h;
            *reinterpret_cast<float*>(tile + lo_off + off) = tf32_rna(vv[r] - h);
        } else {
            *reinterpret_cast<float*>(tile + off) = vv[r];
        }
    }
}

// units [u + i*stride), i < B, of a table, B at a time: all table reads, then all raw reads, then the stores
template <bool X3, int B>
__device__ __forceinline__ void transpose_units(const int4* tab, int u_begin, int u_end, int stride, const uint8_t* raw,
                                                int span, int lane_row, uint8_t* tile, int lo_off, int lane) {
    for (int u = u_begin; u < u_end; u += stride * B) {
        int4 e[B];
        float4 v[B];
#pragma unroll
        for (int i = 0; i < B; ++i) e[i] = tab[min(u + i * stride, u_end - 1)];
#pragma unroll
        for (int i = 0; i < B; ++i) {
            const int src_row = lane_row + e[i].z;
            const int x = ((src_row * span) >> 7) & ((span >> 4) - 1);
            v[i] = *reinterpret_cast<const float4*>(raw + e[i].x + src_row * span + ((e[i].y ^ x) << 4));
        }
        const uint32_t lu = (uint32_t)(lane >> 2), lb = (uint32_t)((lane & 3) << 2);
#pragma unroll
        for (int i = 0; i < B; ++i) {
            if (u + i * stride < u_end) {
                const float vv[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const uint32_t row = (uint32_t)(e[i].w + r);
                    const uint32_t off = row * 128u + (((lu ^ (row & 7u)) << 4) | lb);
                    if (X3) {
                        // hi = the raw value (kind::tf32 reads its top 19 bits), lo = v - trunc(v): exact, 2 ALU ops
                        *reinterpret_cast<float*>(tile + off) = vv[r];
                        *reinterpret_cast<float*>(tile + lo_off + off) = tf32_lo_of_trunc(vv[r]);
                    } else {
                        *reinterpret_cast<float*>(tile + off) = vv[r];
                    }
                }
            }
        }
    }
}

template <bool X3>
__global__ void __launch_bounds__(kWg2Threads, 1)
conv_tc_wgrad2_kernel(const __grid_constant__ CUtensorMap tmap_p, const __grid_constant__ CUtensorMap tmap_q,
                      const Wg2Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_rfull[kWg2MaxStages];
    __shared__ __align__(8) uint64_t bar_rfree[kWg2MaxStages];
    __shared__ __align__(8) uint64_t bar_afull[kWg2MaxStages];
    __shared__ __align__(8) uint64_t bar_aempty[kWg2MaxStages];
    __shared__ __align__(8) uint64_t bar_qfull[2];
    __shared__ __align__(8) uint64_t bar_qempty[2];
    __shared__ __align__(8) uint64_t bar_accum;
    __shared__ uint32_t tmem_base_smem;
    // {raw box byte offset, 16-byte group in the box row, row offset of the tap inside the halo box, first tile row}
    __shared__ int4 a_tab[kWg2MaxAUnits];
    __shared__ int4 q_tab[kWg2MaxQUnits];

    // warp index through a shuffle: ptxas then treats it (and the role branches) as warp-uniform, see conv_tc_halo.cu
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    WG2_KSTAMP(0);
    pdl_launch_dependents();        // programmatic dependent launch: see common.cuh

    // role: input-channel group x output-channel block x row group
    const int rg = blockIdx.y % p.nrg;
    const int cob = (blockIdx.y / p.nrg) % p.ncob;
    const int cig = blockIdx.y / (p.nrg * p.ncob);
    const int qwarps = p.qwarps, awarps = kWg2TransWarps - p.qwarps;
    const int ca0 = cig * p.CaG;
    const int ca_n = min(p.CaG, p.Ca - ca0);
    const int cb0 = cob * p.Nb;
    const int cb_n = min(p.Nb, p.Cb - cb0);
    const int Nmma = (cb_n + 15) & ~15;
    const int taps = p.KH * p.KW;
    const int rows = taps * ca_n;                 // stacked P^T rows of this input-channel group
    const int blk0 = rg * p.bpr;                  // first block of this role
    const int nblk = min(p.bpr, (rows + p.BR - 1) / p.BR - blk0);
    if (nblk <= 0) return;
    const int upb = p.BR >> 2;                    // units per block
    const int n_aunits = rows >> 2;
    const int n_qunits = cb_n >> 2;
    const int nbox_p = ca_n / p.kc_p;
    const int nbox_q = cb_n / p.kc_q;
    const int t_begin = blockIdx.x * p.tiles_per_split;
    const int my_tiles = min(p.ntiles, t_begin + p.tiles_per_split) - t_begin;
    if (my_tiles <= 0) return;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.rstages; ++s) {
            mbar_init(smem_u32(&bar_rfull[s]), 1);
            mbar_init(smem_u32(&bar_rfree[s]), kWg2TransWarps);
        }
        for (int s = 0; s < p.astages; ++s) {
            mbar_init(smem_u32(&bar_afull[s]), awarps);
            mbar_init(smem_u32(&bar_aempty[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&bar_qfull[s]), qwarps);
            mbar_init(smem_u32(&bar_qempty[s]), 1);
        }
        mbar_init(smem_u32(&bar_accum), 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmap_p);
        tma_prefetch_desc(&tmap_q);
    }
    for (int u = threadIdx.x; u < n_aunits; u += blockDim.x) {
        const int r0 = u << 2;
        const int tap = r0 / ca_n, ca = r0 - tap * ca_n;
        const int kh = tap / p.KW, kw = tap - kh * p.KW;
        const int b = ca / p.kc_p;
        a_tab[u] = make_int4(b * p.box_p, (ca - b * p.kc_p) >> 2, kh * p.PW + kw, r0 % p.BR);
    }
    for (int u = threadIdx.x; u < n_qunits; u += blockDim.x) {
        const int cb = u << 2;
        const int b = cb / p.kc_q;
        q_tab[u] = make_int4(p.nbox_p_max * p.box_p + b * p.box_q, (cb - b * p.kc_q) >> 2, 0, cb);
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_smem;
    pdl_wait();                     // the prologue above used shared memory / TMEM only
    WG2_KSTAMP(1);

    if (warp == 0) {
        // ===================== TMA producer =====================
        uint32_t elected;
        asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(elected));
        const bool leader = elected != 0;
        const uint32_t tx_bytes = (uint32_t)(nbox_p * p.tx_p + nbox_q * p.tx_q);
        const uint32_t rawq = (uint32_t)(p.nbox_p_max * p.box_p);
        for (int it = 0; it < my_tiles; ++it) {
            const int s = it % p.rstages;
            mbar_wait(smem_u32(&bar_rfree[s]), (uint32_t)(((it / p.rstages) & 1) ^ 1));
            WG2_STAMP(0);
            const uint32_t full = smem_u32(&bar_rfull[s]);
            const int tile = t_begin + it;
            const int img = tile / p.tiles_per_img;
            const int trem = tile - img * p.tiles_per_img;
            const int ty = trem / p.tiles_x, tx = trem - ty * p.tiles_x;
            const int y0 = ty * p.BH, x0 = tx * p.BW;
            const uint32_t sp = smem_base + (uint32_t)p.raw_base + (uint32_t)s * (uint32_t)p.raw_bytes;
            if (leader) {
                mbar_arrive_expect_tx(full, tx_bytes);
                for (int b = 0; b < nbox_p; ++b)
                    tma_load_4d(sp + (uint32_t)(b * p.box_p), &tmap_p, full, ca0 + b * p.kc_p, x0 - p.pad_l,
                                y0 - p.pad_t, img);
                for (int b = 0; b < nbox_q; ++b)
                    tma_load_4d(sp + rawq + (uint32_t)(b * p.box_q), &tmap_q, full, cb0 + b * p.kc_q, x0, y0, img);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // the whole warp walks the loops (uniform registers), one elected lane issues; descriptors are two running 64-bit
        // values + compile-time K offsets (conv_tc_halo.cu has the measurements behind this shape)
        uint32_t elected;
        asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(elected));
        const bool leader = elected != 0;
        const uint64_t tmpl = make_smem_desc(0, 16, 1024, kLayoutSw128);
        const uint64_t la16 = (uint64_t)((p.swap ? p.q_half : p.a_half) >> 4), lb16 = (uint64_t)((p.swap ? p.a_half : p.q_half) >> 4);
        int item = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const int qs = it % p.qstages;
            mbar_wait(smem_u32(&bar_qfull[qs]), (uint32_t)((it / p.qstages) & 1));
            WG2_STAMP(1);
            const uint32_t qb = smem_base + (uint32_t)p.q_base + (uint32_t)(qs * p.q_slot);
            for (int b = 0; b < nblk; ++b, ++item) {
                const int as = item % p.astages;
                mbar_wait(smem_u32(&bar_afull[as]), (uint32_t)((item / p.astages) & 1));
                if (b == 0) WG2_STAMP(2);
                tc_fence_after();
                const uint32_t ab = smem_base + (uint32_t)p.a_base + (uint32_t)(as * p.a_slot);
                // swap: D[cb][stacked row] with N = rows of this block; else D[stacked row][cb] with N = Nmma
                const int nb_rows = min(p.BR, rows - (blk0 + b) * p.BR);
                const uint32_t idesc = make_idesc_tf32(128, p.swap ? ((nb_rows + 15) & ~15) : Nmma, 0, 0);
                const uint32_t td = tmem_d + (uint32_t)(b * (p.swap ? p.BR : Nmma));
                const uint32_t sa = p.swap ? qb : ab, sb = p.swap ? ab : qb;
                const uint64_t da = tmpl + (uint64_t)((sa & 0x3FFFFu) >> 4);
                const uint64_t db = tmpl + (uint64_t)((sb & 0x3FFFFu) >> 4);
                const uint32_t acc0 = it > 0 ? 1u : 0u;
                if (leader) {
                    if (X3 && p.stackm) {
                        // the 128 A rows starting at the slot hold [Q_hi ; Q_lo]: rows [0,q_rows) of D collect
                        // hi*hi + hi*lo, rows [q_rows, 2 q_rows) lo*hi + lo*lo; the epilogue adds both row groups
                        // to the same dw element.  2 MMAs instead of 3 per K-step.
                        umma_tf32(td, da, db, idesc, acc0);
                        umma_tf32(td, da, db + lb16, idesc, 1u);
#pragma unroll
                        for (int k = 1; k < 4; ++k) {
                            umma_tf32(td, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1u);
                            umma_tf32(td, da + (uint64_t)(2 * k), db + lb16 + (uint64_t)(2 * k), idesc, 1u);
                        }
                    } else if (X3) {
                        umma_tf32(td, da + la16, db, idesc, acc0);
                        umma_tf32(td, da, db + lb16, idesc, 1u);
                        umma_tf32(td, da, db, idesc, 1u);
#pragma unroll
                        for (int k = 1; k < 4; ++k) {
                            umma_tf32(td, da + la16 + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1u);
                            umma_tf32(td, da + (uint64_t)(2 * k), db + lb16 + (uint64_t)(2 * k), idesc, 1u);
                            umma_tf32(td, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1u);
                        }
                    } else {
                        umma_tf32(td, da, db, idesc, acc0);
#pragma unroll
                        for (int k = 1; k < 4; ++k) umma_tf32(td, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1u);
                    }
                    umma_commit(smem_u32(&bar_aempty[as]));
                }
            }
            if (leader) umma_commit(smem_u32(&bar_qempty[qs]));
            WG2_STAMP(3);
        }
        if (leader) umma_commit(smem_u32(&bar_accum));
        __syncwarp();
    } else {
        // ===================== transposers (+ tf32 split), then the epilogue =====================
        const int tw = warp - 2;
        if (tw < qwarps) {
            // ---- Q group: Q^T of every chunk into the Q slot ring
            for (int it = 0; it < my_tiles; ++it) {
                const int rs = it % p.rstages;
                const int qs = it % p.qstages;
                mbar_wait(smem_u32(&bar_rfull[rs]), (uint32_t)((it / p.rstages) & 1));
                if (tw == 0) WG2_STAMP(4);
                mbar_wait(smem_u32(&bar_qempty[qs]), (uint32_t)(((it / p.qstages) & 1) ^ 1));
                if (tw == 0) WG2_STAMP(5);
                transpose_units<X3, 3>(q_tab, tw, n_qunits, qwarps, smem_al + p.raw_base + (size_t)rs * p.raw_bytes, p.span_q,
                                       lane, smem_al + p.q_base + (size_t)qs * p.q_slot, p.q_half, lane);
                if (tw == 0) WG2_STAMP(6);
                fence_proxy_async_smem();
                mbar_arrive_warp(smem_u32(&bar_qfull[qs]));
                mbar_arrive_warp(smem_u32(&bar_rfree[rs]));
                if (tw == 0) WG2_STAMP(7);
            }
        } else {
            // ---- A group: every block of the stacked P^T of this role, through the ring of operand slots
            const int gw = tw - qwarps;
            const int ry = lane / p.BW, rx = lane - ry * p.BW;
            const int lane_row_p = ry * p.PW + rx;          // this lane's pixel inside the halo box (before the tap shift)
            int item = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int rs = it % p.rstages;
                mbar_wait(smem_u32(&bar_rfull[rs]), (uint32_t)((it / p.rstages) & 1));
                const uint8_t* raw = smem_al + p.raw_base + (size_t)rs * p.raw_bytes;
                for (int b = 0; b < nblk; ++b, ++item) {
                    const int as = item % p.astages;
                    mbar_wait(smem_u32(&bar_aempty[as]), (uint32_t)(((item / p.astages) & 1) ^ 1));
                    if (gw == 0 && b == 0) WG2_STAMP(8);
                    transpose_units<X3, 3>(a_tab, (blk0 + b) * upb + gw, min(n_aunits, (blk0 + b + 1) * upb), awarps, raw,
                                           p.span_p, lane_row_p, smem_al + p.a_base + (size_t)as * p.a_slot, p.a_half, lane);
                    if (gw == 0 && b == 0) WG2_STAMP(9);
                    fence_proxy_async_smem();
                    mbar_arrive_warp(smem_u32(&bar_afull[as]));
                    if (gw == 0 && b == 0) WG2_STAMP(10);
                }
                mbar_arrive_warp(smem_u32(&bar_rfree[rs]));
                if (gw == 0) WG2_STAMP(11);
            }
        }
        {
            // ---- epilogue: TMEM lane quadrant = warp % 4; the 4 warps of a quadrant take alternate 16-column groups
            mbar_wait(smem_u32(&bar_accum), 0);
            tc_fence_after();
            WG2_KSTAMP(2);
            const int q = warp & 3;
            const int cphase = tw >> 2;
            if (!p.swap) {
                for (int b = 0; b < nblk; ++b) {
                    const int row = ((blk0 + b) << 7) + q * 32 + lane;   // stacked row = accumulator lane
                    const bool row_ok = row < rows;
                    const int tap = row / ca_n, ca = row - tap * ca_n;
                    float* dst_row = p.dw + ((int64_t)tap * p.Ca + ca0 + ca) * p.Cb + cb0;
                    const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * Nmma);
                    for (int c0 = cphase * 16; c0 < Nmma; c0 += 64) {
                        float v[16];
                        tmem_ld16(taddr + (uint32_t)c0, v);
                        if (row_ok) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                if (c0 + j < cb_n) red_add_v4_f32(dst_row + c0 + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
                        }
                    }
                }
            } else {
                // accumulator lane = output channel (stackm: a second group of lanes holds the lo-half products of
                // the same channels), column = stacked row.  Staged through shared memory (free now: every TMA box
                // is consumed and every MMA retired) so that the hi/lo lane groups are summed on chip and dw gets
                // 16-byte vector reductions along cb -- per-lane scalar reductions of both groups were ~40 % of
                // this kernel's time on the backbone layers (L2 atomic throughput, 148 partial tiles).
                float* const T = reinterpret_cast<float*>(smem_al);
                const int tp = p.stackm ? 2 * p.q_rows : p.q_rows;       // tile pitch (floats), multiple of 8
                const int r = q * 32 + lane;
                const int cb4n = cb_n >> 2;
                const int et = threadIdx.x - 64;                          // 0..511
                for (int b = 0; b < nblk; ++b) {
                    const int r_base = (blk0 + b) * p.BR;
                    const int nb_rows = min(p.BR, rows - r_base);
                    if (q * 32 < tp) {
                        const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * p.BR);
                        for (int c0 = cphase * 16; c0 < nb_rows; c0 += 64) {
                            float v[16];
                            tmem_ld16(taddr + (uint32_t)c0, v);
                            if (r < tp) {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    if (c0 + j < nb_rows) T[(c0 + j) * tp + r] = v[j];
                            }
                        }
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(kWg2TransWarps * 32) : "memory");
                    for (int idx = et; idx < nb_rows * cb4n; idx += kWg2TransWarps * 32) {
                        const int j = idx / cb4n, c4 = (idx - j * cb4n) << 2;
                        float4 sum = *reinterpret_cast<const float4*>(T + j * tp + c4);
                        if (p.stackm) {
                            const float4 lo = *reinterpret_cast<const float4*>(T + j * tp + p.q_rows + c4);
                            sum.x += lo.x; sum.y += lo.y; sum.z += lo.z; sum.w += lo.w;
                        }
                        const int row = r_base + j;
                        const int tap = row / ca_n, ca = row - tap * ca_n;
                        red_add_v4_f32(p.dw + ((int64_t)tap * p.Ca + ca0 + ca) * p.Cb + cb0 + c4, sum.x, sum.y, sum.z, sum.w);
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(kWg2TransWarps * 32) : "memory");
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    WG2_KSTAMP(3);
    if (warp == 1) tmem_dealloc(tmem_d, (uint32_t)p.tmem_cols);
}

// -------------------------------------------------------------------------------------------------
// host dispatch
// -------------------------------------------------------------------------------------------------
static bool chunk_geometry(int H, int W, int* BW, int* BH) {
    int bw;
    if (W >= 32) {
        if (W % 32) return false;
        bw = 32;
    } else {
        if (W < 8 || 32 % W) return false;
        bw = W;
    }
    const int bh = 32 / bw;
    if (H % bh) return false;
    *BW = bw;
    *BH = bh;
    return true;
}

static long long* g_wg2_dbg = nullptr;
void wgrad2_set_debug_buffer(long long* p) { g_wg2_dbg = p; }

int conv2d_wgrad_tc2(const WgradArgs& a, int math_mode, cudaStream_t st) {
    static const bool disabled = [] { const char* e = getenv("DL4DS_WGRAD_V1"); return e && e[0] == '1'; }();
    if (disabled) return DL4DS_E_UNSUPPORTED;
    if (math_mode != DL4DS_MATH_TF32 && math_mode != DL4DS_MATH_TF32X3) return DL4DS_E_UNSUPPORTED;
    if (dl4ds_device_is_sm100() != 1) return DL4DS_E_UNSUPPORTED;
    if (a.stride != 1 || a.Hp != a.Hq || a.Wp != a.Wq) return DL4DS_E_UNSUPPORTED;
    if (a.Ca % 8 || a.Cb % 8 || a.p_ld % 4 || a.q_ld % 4) return DL4DS_E_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(a.P) & 15) || (reinterpret_cast<uintptr_t>(a.Q) & 15) ||
        (reinterpret_cast<uintptr_t>(a.dw) & 15))
        return DL4DS_E_UNSUPPORTED;
    if (a.KH > 7 || a.KW > 7) return DL4DS_E_UNSUPPORTED;      // (7x7: the ConvNeXt stem / tail, validated in round 2)
    if (a.KH * a.KW == 1) return DL4DS_E_UNSUPPORTED;     // 1x1: a handful of stacked rows -- the first-generation kernel is faster (measured)
    const bool x3 = math_mode == DL4DS_MATH_TF32X3;
    Wg2Params p;
    if (!chunk_geometry(a.Hq, a.Wq, &p.BW, &p.BH)) return DL4DS_E_UNSUPPORTED;
    const Chunk cp = pick_chunk(a.Ca), cq = pick_chunk(a.Cb);
    const int taps = a.KH * a.KW;
    p.dw = a.dw;
    p.dbg = g_wg2_dbg;
    p.H = a.Hq; p.W = a.Wq; p.Ca = a.Ca; p.Cb = a.Cb;
    p.KH = a.KH; p.KW = a.KW; p.pad_t = a.pad_t; p.pad_l = a.pad_l;
    p.PW = p.BW + a.KW - 1;
    const int PH = p.BH + a.KH - 1;
    p.tiles_x = a.Wq / p.BW;
    p.tiles_per_img = p.tiles_x * (a.Hq / p.BH);
    p.ntiles = a.N * p.tiles_per_img;
    p.kc_p = cp.kc; p.span_p = cp.span;
    p.kc_q = cq.kc; p.span_q = cq.span;
    // input-channel group: all of Ca when the stacked rows fit 512 accumulator rows/columns, else the largest
    // multiple of kc that does
    int CaG = a.Ca;
    if (taps * CaG > 512) CaG = (512 / taps) / cp.kc * cp.kc;
    if (CaG < cp.kc) return DL4DS_E_UNSUPPORTED;
    p.CaG = CaG;
    p.ncig = (a.Ca + CaG - 1) / CaG;
    const int rows_g = taps * CaG;
    const int unit = cq.kc > 16 ? cq.kc : 16;
    p.swap = a.Cb <= 128 ? 1 : 0;
    int nmma_max, cb_first, tmem_need;
    if (p.swap) {
        // A = Q^T (all of Cb, one M tile), B = blocks of <= 256 stacked rows, all blocks in one role
        p.Nb = a.Cb; p.ncob = 1;
        cb_first = a.Cb;
        nmma_max = (a.Cb + 15) & ~15;
        // block height (DL4DS_WG2_BR_MAX, default 256).  Smaller blocks would let the A-slot ring turn over inside a
        // chunk, but every block re-reads the Q^T operand: measured (round 2, 48 -> 48 @ 32x32) 4050 clk per chunk with
        // two 216-row blocks against 5100 / 6000 clk with four 112-row / five 96-row blocks -- the kernel is bound by
        // shared-memory operand bandwidth, not by the MMA count
        static const int br_max = [] { const char* e = getenv("DL4DS_WG2_BR_MAX"); int v = e ? atoi(e) : 256; return v < 64 ? 64 : (v > 256 ? 256 : v); }();
        int nnb = (rows_g + br_max - 1) / br_max;
        while (nnb * ((((rows_g + nnb - 1) / nnb) + 15) & ~15) > 512) --nnb;      // accumulators: nnb * BR TMEM columns
        if (nnb < (rows_g + 255) / 256) nnb = (rows_g + 255) / 256;
        p.BR = ((rows_g + nnb - 1) / nnb + 15) & ~15;
        p.bpr = nnb; p.nrg = 1;
        // one block per CTA role (DL4DS_WG2_SPLIT_ROLES=1): each role then keeps several CHUNKS of its block in the
        // A-slot ring, so the transposer of chunk i + 1 overlaps the MMAs of chunk i (with all blocks in one role the
        // slots that fit hold exactly one chunk and the two phases alternate)
        static const int split_roles = [] { const char* e = getenv("DL4DS_WG2_SPLIT_ROLES"); return e ? atoi(e) : 0; }();
        if (split_roles && nnb > 1) { p.bpr = 1; p.nrg = nnb; }
        tmem_need = p.bpr * p.BR;
    } else {
        // A = 128-row blocks of stacked rows, B = Q^T with N = Cb (<= 256 per output-channel role); as many blocks
        // per role as accumulators fit TMEM, the rest goes to further row-group roles
        int ncob = (a.Cb + 255) / 256;
        int nb = ((a.Cb + ncob - 1) / ncob + unit - 1) / unit * unit;
        if (nb > 256) { nb = 256 / unit * unit; ncob = (a.Cb + nb - 1) / nb; }
        p.Nb = nb; p.ncob = ncob;
        cb_first = a.Cb < nb ? a.Cb : nb;
        nmma_max = (cb_first + 15) & ~15;
        p.BR = 128;
        const int nblk_total = (rows_g + 127) / 128;
        int bpr = 512 / nmma_max;
        if (bpr > nblk_total) bpr = nblk_total;
        p.bpr = bpr;
        p.nrg = (nblk_total + bpr - 1) / bpr;
        tmem_need = bpr * nmma_max;
    }
    if (rows_g / 4 > kWg2MaxAUnits || cb_first / 4 > kWg2MaxQUnits) return DL4DS_E_UNSUPPORTED;
    p.tx_p = cp.span * p.PW * PH;
    p.tx_q = cq.span * 32;
    p.box_p = (p.tx_p + 1023) & ~1023;
    p.box_q = (p.tx_q + 1023) & ~1023;
    p.nbox_p_max = CaG / cp.kc;
    p.raw_bytes = p.nbox_p_max * p.box_p + (cb_first / cq.kc) * p.box_q;
    p.a_half = p.BR * 128;
    p.q_rows = p.swap ? ((cb_first + 7) & ~7) : nmma_max;
    p.q_half = p.q_rows * 128;
    static const bool no_stack = [] { const char* e = getenv("DL4DS_TC_NO_STACKM"); return e && e[0] == '1'; }();
    p.stackm = (p.swap && x3 && 2 * p.q_rows <= 128 && !no_stack) ? 1 : 0;
    p.a_slot = p.a_half * (x3 ? 2 : 1);
    p.q_slot = p.q_half * (x3 ? 2 : 1);
    int cols = 32;
    while (cols < tmem_need) cols *= 2;
    if (cols > 512) return DL4DS_E_UNSUPPORTED;
    p.tmem_cols = cols;
    // transposer warps: the Q group gets a share proportional to its units per chunk (4 or 8 of the 16 warps)
    const int a_units_chunk = (p.bpr * p.BR < rows_g ? p.bpr * p.BR : rows_g) / 4;
    p.qwarps = (cb_first / 4) * 3 >= a_units_chunk * 2 ? 8 : 4;
    // shared-memory plan: [Q slots][raw TMA ring][A-slot ring].  Minimum 1 Q slot, 2 raw stages, 2 A slots; spare
    // capacity goes to a 2nd Q slot, a 3rd raw stage, a 3rd A slot, then raw stages up to 6.  (swap: an A-operand
    // read of the short Q^T tile covers 128 rows = 16 KB from the slot start and runs on into the raw ring.)
    const int budget = 220 * 1024;
    auto need = [&](int qn, int r, int s) { return qn * p.q_slot + r * p.raw_bytes + s * p.a_slot; };
    int qst = 1, rst = 2, ast = 2;
    if (need(qst, rst, ast) > budget) return DL4DS_E_UNSUPPORTED;
    if (need(2, rst, ast) <= budget) qst = 2;
    if (need(qst, 3, ast) <= budget) rst = 3;
    if (need(qst, rst, 3) <= budget) ast = 3;
    while (ast < kWg2MaxStages && ast < (p.bpr == 1 ? 4 : 2 * p.bpr) && need(qst, rst, ast + 1) <= budget) ++ast;
    while (rst < 6 && need(qst, rst + 1, ast) <= budget) ++rst;
    p.qstages = qst; p.rstages = rst; p.astages = ast;
    p.q_base = 0;
    p.raw_base = qst * p.q_slot;
    p.a_base = p.raw_base + rst * p.raw_bytes;
    size_t smem = (size_t)need(qst, rst, ast) + 1024;
    if (p.swap && smem < (size_t)(p.q_slot * qst + 33 * 1024)) smem = (size_t)(p.q_slot * qst + 33 * 1024);
    const size_t epi_tile = (size_t)p.BR * (p.stackm ? 2 * p.q_rows : p.q_rows) * 4 + 1024;   // swap epilogue staging
    if (p.swap && smem < epi_tile) smem = epi_tile;
    if (smem > 221 * 1024) return DL4DS_E_UNSUPPORTED;
    const int nroles = p.ncig * p.ncob * p.nrg;
    int splits = kNumSMs / nroles;
    if (splits < 1) splits = 1;
    if (splits > p.ntiles) splits = p.ntiles;
    p.tiles_per_split = (p.ntiles + splits - 1) / splits;
    splits = (p.ntiles + p.tiles_per_split - 1) / p.tiles_per_split;
    const CUtensorMap* tp = get_tensor_map_nhwc(a.P, a.p_ld, a.N, a.Hp, a.Wp, a.Ca, cp.kc, p.PW, PH, cp.swz);
    const CUtensorMap* tq = get_tensor_map_nhwc(a.Q, a.q_ld, a.N, a.Hq, a.Wq, a.Cb, cq.kc, p.BW, p.BH, cq.swz);
    if (!tp || !tq) return DL4DS_E_CUDA;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(conv_tc_wgrad2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(222 * 1024));
        cudaFuncSetAttribute(conv_tc_wgrad2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(222 * 1024));
        attr_done = true;
    }
    dim3 grid((unsigned)splits, (unsigned)nroles);
    if (x3)
        launch_pdl(2, conv_tc_wgrad2_kernel<true>, grid, dim3(kWg2Threads), smem, st, *tp, *tq, p);
    else
        launch_pdl(2, conv_tc_wgrad2_kernel<false>, grid, dim3(kWg2Threads), smem, st, *tp, *tq, p);
    g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    return check_launch("conv_tc_wgrad2_kernel");
}

}  // namespace dl4ds
