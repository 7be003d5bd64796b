// Probe (round 2): does the halo kernel's operand pattern cost more than the dense SWIZZLE_128B tiles of umma_rate3?
//   pattern 0: dense SW128 tiles (SBO 1024), one N                          -> reference (44-56 clk)
//   pattern 1: dense SW64 tiles (rows 64 B, SBO 512)
//   pattern 2: SW64 halo view: SBO 640 (10-pixel pitch), tap-shifted start rows
//   pattern 3: pattern 2 + alternating N=2*Npad / N=Npad MMAs (the stacked 3xTF32 pair), A_hi / A_lo tiles
//   pattern 4: dense SW128 + alternating N
//   pattern 5: SW128 halo view (rows 128 B, SBO 1280), alternating N
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../dl4ds_b200/csrc/tc_common.cuh"
using namespace dl4ds::tc;
namespace dl4ds { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

__device__ __forceinline__ void umma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc) : "memory");
}
struct Res { long long issue, complete; };

template <int PATTERN>
__global__ void rate(Res* out, int Npad, int iters8) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint32_t* sm = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < 120 * 1024 / 4; i += blockDim.x) sm[i] = 0x3F800000u;
    fence_proxy_async_smem();
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t td = slot;
    if (warp == 1) {
        uint32_t pred;
        asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
        const bool leader = pred != 0;
        const bool alt = PATTERN == 3 || PATTERN == 4 || PATTERN == 5;
        const uint32_t idesc1 = make_idesc_tf32(128, alt ? 2 * Npad : Npad, 0, 0);
        const uint32_t idesc2 = make_idesc_tf32(128, Npad, 0, 0);
        const int span = (PATTERN == 0 || PATTERN == 4 || PATTERN == 5) ? 128 : 64;
        const uint32_t layout = span == 128 ? kLayoutSw128 : kLayoutSw64;
        const uint32_t sbo_a = (PATTERN == 2 || PATTERN == 3 || PATTERN == 5) ? 10u * span : 8u * span;
        const uint64_t ta = make_smem_desc(0, 16, sbo_a, layout);
        const uint64_t tb = make_smem_desc(0, 16, 8u * span, layout);
        uint64_t da[8], db[8];
        uint32_t id[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t aoff, boff;
            if (PATTERN == 2 || PATTERN == 3 || PATTERN == 5) {
                // k-step pair j>>1 of tap (kh, kw) = ((j>>2)&1, j>>2): start row kh*10 + kw; hi tile / lo tile (+32 KB) alternate
                const int tap = j >> 2, kk = (j >> 1) & 1;
                aoff = (uint32_t)((tap * 11) * span + kk * 32 + ((alt && (j & 1)) ? 32768 : 0));
                boff = (uint32_t)(tap * 2 * Npad * span + kk * 32);
            } else {
                aoff = (uint32_t)((j >> 2) * 16384 + (j & 3 & (span / 32 - 1)) * 32);
                boff = (uint32_t)((j & 3 & (span / 32 - 1)) * 32);
            }
            da[j] = ta + (uint64_t)(((base + aoff) & 0x3FFFFu) >> 4);
            db[j] = tb + (uint64_t)(((base + 80 * 1024 + boff) & 0x3FFFFu) >> 4);
            id[j] = (alt && (j & 1)) ? idesc2 : idesc1;
        }
        const long long t0 = clock64();
        for (int i = 0; i < iters8; ++i) {
            if (leader) {
#pragma unroll
                for (int j = 0; j < 8; ++j) umma(td, da[j], db[j], id[j]);
            }
        }
        const long long t1 = clock64();
        if (leader) umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        const long long t2 = clock64();
        if (leader) { out[blockIdx.x].issue = t1 - t0; out[blockIdx.x].complete = t2 - t0; }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(td, 512);
}

static Res* d_out;
static Res h_out[256];
template <int PATTERN>
static void run(int Npad) {
    const int iters8 = 512, grid = 148;
    cudaFuncSetAttribute(rate<PATTERN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
    rate<PATTERN><<<grid, 128, 130 * 1024>>>(d_out, Npad, iters8);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(h_out, d_out, sizeof(Res) * grid, cudaMemcpyDeviceToHost);
    double sc = 0, si = 0;
    for (int i = 0; i < grid; ++i) { sc += (double)h_out[i].complete; si += (double)h_out[i].issue; }
    printf("pattern %d Npad=%3d: issue %.1f, complete %.1f clk/MMA\n", PATTERN, Npad, si / (grid * iters8 * 8.0), sc / (grid * iters8 * 8.0));
    fflush(stdout);
}
int main() {
    cudaMalloc(&d_out, sizeof(Res) * 256);
    for (int Npad : {32, 48, 64}) { run<0>(Npad); run<1>(Npad); run<2>(Npad); run<3>(Npad); run<4>(Npad); run<5>(Npad); }
    return 0;
}
