// Probe for the round-2 plan: cycles per tcgen05.mma kind::f16 (fp16 operands, fp32
// accumulate, M=128, K=16 = one 32-byte slice per row, the same operand bytes as a kind::tf32 K=8 MMA) as a function
// of N, next to kind::tf32 -- an fp16 3-term split would issue half as many MMAs per MAC as 3xTF32 if the
// per-instruction floor is the same.  Same harness as umma_rate.cu (SS mode, one accumulator, 1 CTA).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../dl4ds_b200/csrc/tc_common.cuh"
using namespace dl4ds::tc;
namespace dl4ds { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

__host__ __device__ inline uint32_t make_idesc_f16(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                       // c_format = F32
    d |= 0u << 7;                       // a_format = F16
    d |= 0u << 10;                      // b_format = F16
    d |= static_cast<uint32_t>(N >> 3) << 17;
    d |= static_cast<uint32_t>(M >> 4) << 24;
    return d;
}

__global__ void rate(long long* out, int N, int kind, int iters) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint32_t* sm = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)));
    // fp16 1.0 pairs (0x3C003C00) or fp32 1.0f: finite operands either way
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) sm[i] = kind ? 0x3C003C00u : 0x3F800000u;
    fence_proxy_async_smem();
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t td = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = kind ? make_idesc_f16(128, N) : make_idesc_tf32(128, N, 0, 0);
        const uint32_t a_s = base, b_s = base + 16 * 1024;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t ko = (uint32_t)(i & 3) * 32u;
            const uint64_t da = make_smem_desc(a_s + ko, 16, 1024, kLayoutSw128);
            const uint64_t db = make_smem_desc(b_s + ko, 16, 1024, kLayoutSw128);
            if (kind) umma_f16(td, da, db, idesc, 1u);
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
    const char* names[2] = {"tf32 K=8 ", "f16  K=16"};
    const int iters = 4096;
    for (int kind = 0; kind < 2; ++kind)
        for (int N : {16, 32, 48, 64, 96, 128, 192, 256}) {
            rate<<<1, 128, 50 * 1024>>>(d, N, kind, iters);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%s N=%3d: complete %.1f clk/MMA = %.0f MAC/clk\n", names[kind], N, (double)h[1] / iters,
                   128.0 * N * (kind ? 16 : 8) / ((double)h[1] / iters));
        }
    return 0;
}
