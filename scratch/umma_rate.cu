// Probe: cycles per tcgen05.mma kind::tf32 (M=128, K=8) as a function of N, for
//   SS  = A and B from shared memory, one accumulator        (what conv_tc*.cu issue today)
//   SS2 = same, alternating between two accumulators         (is the fixed cost a D dependency?)
//   TS  = A from TMEM, B from shared memory                  (is the fixed cost the A-operand smem read?)
//   SSk = SS, but all MMAs read the same 32-byte K slice     (operand address pattern)
// One CTA per SM is launched on `ctas` SMs so that nothing else competes for the SM's shared memory.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../dl4ds_b200/csrc/tc_common.cuh"
using namespace dl4ds::tc;
namespace dl4ds { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

__global__ void rate(long long* out, int N, int mode, int iters) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    float* sm = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) sm[i] = 1.0f;     // A: 16 KB tile, B: up to 32 KB
    fence_proxy_async_smem();
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t td = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_tf32(128, N, 0, 0);
        const uint32_t a_s = base, b_s = base + 16 * 1024;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t ko = (mode == 3) ? 0u : (uint32_t)(i & 3) * 32u;
            const uint64_t da = make_smem_desc(a_s + ko, 16, 1024, kLayoutSw128);
            const uint64_t db = make_smem_desc(b_s + ko, 16, 1024, kLayoutSw128);
            if (mode == 2) umma_ts(td, td + 256 + (i & 3) * 8, db, idesc, 1u);
            else if (mode == 1) umma_tf32(td + (uint32_t)((i & 1) * 256), da, db, idesc, 1u);
            else umma_tf32(td, da, db, idesc, 1u);
        }
        const long long t1 = clock64();
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(td, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const char* names[4] = {"SS ", "SS2", "TS ", "SSk"};
    const int iters = 4096;
    for (int mode = 0; mode < 4; ++mode)
        for (int N : {16, 32, 48, 64, 96, 128, 192, 256}) {
            if (mode == 1 && N > 256) continue;
            rate<<<1, 128, 50 * 1024>>>(d, N, mode, iters);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%s N=%3d: issue %.1f clk/MMA, complete %.1f clk/MMA (ideal %d)\n", names[mode], N,
                   (double)h[0] / iters, (double)h[1] / iters, N);
        }
    return 0;
}
