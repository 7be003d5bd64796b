// Probe (round 2, VERDICT r01 "weak" item 4): is the ~95-cycle cost of a small-N tcgen05.mma that
// scratch/umma_rate.cu measured a THROUGHPUT floor or the latency of a dependent accumulate chain?
//   nacc   = 1 / 2 / 4 independent TMEM accumulators used round-robin
//   cps    = 1 / 2 CTAs per SM issuing at the same time (grid = 148 * cps; per-CTA and per-SM rate reported)
//   kind   = tf32 (K = 8 per MMA) / f16 (K = 16): both read 32 bytes per operand row
//   span   = 128 / 64 / 32-byte operand rows (SWIZZLE_128B / 64B / 32B K-major tiles: what kc = 32 / 16 / 8 layers use)
//   cg2    = cta_group::2 (M = 256 over a CTA pair, every CTA holds half of B)
// Usage: umma_rate2 [cg2]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "../dl4ds_b200/csrc/tc_common.cuh"
using namespace dl4ds::tc;
namespace dl4ds { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f16_cg2(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_tf32_cg2(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ inline uint32_t make_idesc(int kind, int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                                   // c_format = F32
    if (!kind) { d |= 2u << 7; d |= 2u << 10; }     // TF32 operands (F16 = 0)
    d |= static_cast<uint32_t>(N >> 3) << 17;
    d |= static_cast<uint32_t>(M >> 4) << 24;
    return d;
}

struct Res { long long issue, complete; };

__global__ void rate(Res* out, int N, int kind, int nacc, int span, int iters, int tmem_cols) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint32_t* sm = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < 80 * 1024 / 4; i += blockDim.x) sm[i] = kind ? 0x3C003C00u : 0x3F800000u;
    fence_proxy_async_smem();
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&slot), (uint32_t)tmem_cols);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t td = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc(kind, 128, N);
        const uint32_t layout = span == 128 ? kLayoutSw128 : (span == 64 ? kLayoutSw64 : kLayoutSw32);
        const uint32_t sbo = 8u * (uint32_t)span;
        const int ksteps = span / 32;                       // 32-byte K slices per operand row
        // A: 3 tiles of 128 rows (<= 48 KB), B: 256 rows (<= 32 KB) behind them
        const uint32_t a_s = base, b_s = base + 48 * 1024;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t ko = (uint32_t)(i % ksteps) * 32u;
            const uint32_t at = (uint32_t)((i / ksteps) % 3) * (uint32_t)(128 * span);
            const uint64_t da = make_smem_desc(a_s + at + ko, 16, sbo, layout);
            const uint64_t db = make_smem_desc(b_s + ko, 16, sbo, layout);
            const uint32_t d = td + (uint32_t)((i % nacc) * N);
            if (kind) umma_f16(d, da, db, idesc, 1u);
            else umma_tf32(d, da, db, idesc, 1u);
        }
        const long long t1 = clock64();
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        const long long t2 = clock64();
        out[blockIdx.x].issue = t1 - t0;
        out[blockIdx.x].complete = t2 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(td, (uint32_t)tmem_cols);
}

// cta_group::2: a CTA pair, M = 256 (128 rows of A per CTA), each CTA holds N/2 rows of B at the same smem offset.
__global__ void __cluster_dims__(2, 1, 1) rate_cg2(Res* out, int N, int kind, int nacc, int iters) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint32_t* sm = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < 80 * 1024 / 4; i += blockDim.x) sm[i] = kind ? 0x3C003C00u : 0x3F800000u;
    fence_proxy_async_smem();
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    tc_fence_after();
    const uint32_t td = slot;
    if (threadIdx.x == 0 && rank == 0) {
        const uint32_t idesc = make_idesc(kind, 256, N);
        const uint32_t a_s = base, b_s = base + 48 * 1024;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t ko = (uint32_t)(i & 3) * 32u;
            const uint32_t at = (uint32_t)((i >> 2) % 3) * (128u * 128u);
            const uint64_t da = make_smem_desc(a_s + at + ko, 16, 1024, kLayoutSw128);
            const uint64_t db = make_smem_desc(b_s + ko, 16, 1024, kLayoutSw128);
            const uint32_t d = td + (uint32_t)((i % nacc) * N);
            if (kind) umma_f16_cg2(d, da, db, idesc, 1u);
            else umma_tf32_cg2(d, da, db, idesc, 1u);
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(smem_u32(&bar), 0);
        const long long t2 = clock64();
        out[blockIdx.x >> 1].issue = t1 - t0;
        out[blockIdx.x >> 1].complete = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(td), "r"(512u) : "memory");
}

static Res* d_out;
static Res h_out[1024];

static double mean_complete(int n) {
    cudaMemcpy(h_out, d_out, sizeof(Res) * n, cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < n; ++i) s += (double)h_out[i].complete;
    return s / n;
}

int main(int argc, char** argv) {
    cudaMalloc(&d_out, sizeof(Res) * 1024);
    const int iters = 4096;
    const char* kn[2] = {"tf32 K=8 ", "f16  K=16"};
    if (argc > 1 && !strcmp(argv[1], "cg2")) {
        cudaFuncSetAttribute(rate_cg2, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        for (int kind = 0; kind < 2; ++kind)
            for (int nacc : {1, 2})
                for (int N : {32, 64, 96, 128, 256}) {
                    if (nacc * N > 512) continue;
                    rate_cg2<<<2, 128, 90 * 1024>>>(d_out, N, kind, nacc, iters);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
                    const double c = mean_complete(1) / iters;
                    printf("cg2 %s M=256 N=%3d nacc=%d: %.1f clk/MMA = %.0f MAC/clk/SM\n", kn[kind], N, nacc, c,
                           256.0 * N * (kind ? 16 : 8) / c / 2);
                    fflush(stdout);
                }
        return 0;
    }
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int kind = 0; kind < 2; ++kind)
        for (int span : {128, 64, 32})
            for (int cps : {1, 2})
                for (int nacc : {1, 2, 4})
                    for (int N : {32, 48, 64, 96, 128, 192, 256}) {
                        const int cols_needed = nacc * N;
                        const int cap = cps == 2 ? 256 : 512;
                        if (cols_needed > cap) continue;
                        if (span != 128 && (N > 128 || nacc == 4)) continue;       // keep the sweep short
                        int cols = 32;
                        while (cols < cols_needed) cols *= 2;
                        const int grid = 148 * cps;
                        rate<<<grid, 128, 90 * 1024>>>(d_out, N, kind, nacc, span, iters, cps == 2 ? 256 : cols);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
                        const double c = mean_complete(grid) / iters;          // cycles per MMA seen by one CTA
                        printf("%s span=%3d cps=%d nacc=%d N=%3d: %.1f clk/MMA/CTA, %.1f clk/MMA/SM = %.0f MAC/clk/SM\n",
                               kn[kind], span, cps, nacc, N, c, c / cps, 128.0 * N * (kind ? 16 : 8) / (c / cps));
                        fflush(stdout);
                    }
    return 0;
}
