// Probe (round 2): umma_rate2 showed ~249 clk per MMA for EVERY N / kind / accumulator count -- the cost of its own
// scalar issue loop (runtime % and / per iteration), i.e. one thread cannot issue faster than its dependent ALU chain.
// Here the loop is unrolled x8 with every descriptor and accumulator address precomputed in registers, so the
// measured interval is the tensor pipe's (or the operand fetch's), not the issuing thread's.
//   nacc = independent accumulators (round-robin), cps = CTAs per SM, kind tf32 (K=8) / f16 (K=16), N
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../dl4ds_b200/csrc/tc_common.cuh"
using namespace dl4ds::tc;
namespace dl4ds { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

template <int KIND>
__device__ __forceinline__ void umma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc) {
    if (KIND)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(d), "l"(da), "l"(db), "r"(idesc) : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(d), "l"(da), "l"(db), "r"(idesc) : "memory");
}
__host__ __device__ inline uint32_t make_idesc(int kind, int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;
    if (!kind) { d |= 2u << 7; d |= 2u << 10; }
    d |= static_cast<uint32_t>(N >> 3) << 17;
    d |= static_cast<uint32_t>(M >> 4) << 24;
    return d;
}
struct Res { long long issue, complete; };

template <int KIND, int NACC>
__global__ void rate(Res* out, int N, int iters8, int tmem_cols) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint32_t* sm = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < 80 * 1024 / 4; i += blockDim.x) sm[i] = KIND ? 0x3C003C00u : 0x3F800000u;
    fence_proxy_async_smem();
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&slot), (uint32_t)tmem_cols);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t td = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc(KIND, 128, N);
        uint64_t da[8], db[8];
        uint32_t dd[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            // 2 A tiles x 4 K-slices of a 128-byte row; B: 4 K-slices
            da[j] = make_smem_desc(base + (uint32_t)(j >> 2) * 16384u + (uint32_t)(j & 3) * 32u, 16, 1024, kLayoutSw128);
            db[j] = make_smem_desc(base + 48 * 1024 + (uint32_t)(j & 3) * 32u, 16, 1024, kLayoutSw128);
            dd[j] = td + (uint32_t)((j % NACC) * N);
        }
        const long long t0 = clock64();
        for (int i = 0; i < iters8; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) umma<KIND>(dd[j], da[j], db[j], idesc);
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

static Res* d_out;
static Res h_out[1024];

template <int KIND, int NACC>
static void run(int N, int cps) {
    const int iters8 = 512;
    const int cap = cps == 2 ? 256 : 512;
    if (NACC * N > cap) return;
    int cols = 32;
    while (cols < NACC * N) cols *= 2;
    if (cps == 2) cols = 256;
    const int grid = 148 * cps;
    cudaFuncSetAttribute(rate<KIND, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    rate<KIND, NACC><<<grid, 128, 90 * 1024>>>(d_out, N, iters8, cols);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(h_out, d_out, sizeof(Res) * grid, cudaMemcpyDeviceToHost);
    double si = 0, sc = 0;
    for (int i = 0; i < grid; ++i) { si += (double)h_out[i].issue; sc += (double)h_out[i].complete; }
    const double n = (double)grid * iters8 * 8;
    const double ci = si / n, cc = sc / n;
    printf("%s cps=%d nacc=%d N=%3d: issue %.1f, complete %.1f clk/MMA/CTA -> %.1f clk/MMA/SM = %.0f MAC/clk/SM (ideal %d clk)\n",
           KIND ? "f16  K=16" : "tf32 K=8 ", cps, NACC, N, ci, cc, cc / cps, 128.0 * N * (KIND ? 16 : 8) / (cc / cps), N / 2);
    fflush(stdout);
}

int main() {
    cudaMalloc(&d_out, sizeof(Res) * 1024);
    for (int cps : {1, 2})
        for (int N : {16, 32, 48, 64, 96, 128, 192, 256}) {
            run<0, 1>(N, cps); run<0, 2>(N, cps); run<0, 4>(N, cps);
            run<1, 1>(N, cps); run<1, 2>(N, cps); run<1, 4>(N, cps);
        }
    return 0;
}
