// Probe: tcgen05.mma kind::tf32 with A in TMEM (written by tcgen05.st, lane = row m, column = k).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../dl4ds_b200/csrc/tc_common.cuh"
using namespace dl4ds::tc;
namespace dl4ds { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                   "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}

// A: [128][K] row-major in global; B image: K-major SW64 tiles per 16-k chunk: [chunk][N][64B]
__global__ void probe(const float* A, const uint8_t* b_img, float* out, int K, int N, uint32_t b_bytes) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    for (uint32_t i = threadIdx.x; i < b_bytes / 4; i += blockDim.x) ((uint32_t*)sm)[i] = ((const uint32_t*)b_img)[i];
    fence_proxy_async_smem();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t td = slot;            // D at columns [0, N), A at columns [256, 256+K)
    const int row = warp * 32 + lane;
    for (int k0 = 0; k0 < K; k0 += 8) {
        float v[8];
        for (int j = 0; j < 8; ++j) v[j] = A[row * K + k0 + j];
        tmem_st8(td + ((uint32_t)(warp * 32) << 16) + 256 + k0, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_tf32(128, N, 0, 0);
        for (int k0 = 0; k0 < K; k0 += 8) {
            const int chunk = k0 / 16, ko = (k0 % 16) * 4;
            uint64_t db = make_smem_desc(base + chunk * N * 64 + ko, 16, 512, kLayoutSw64);
            umma_tf32_ts(td, td + 256 + k0, db, idesc, k0 > 0);
        }
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(td + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int j = 0; j < 16; ++j) out[row * N + c0 + j] = v[j];
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(td, 512);
}

int main() {
    const int K = 32, N = 48;
    std::vector<float> A(128 * K), B(N * K);
    srand(3);
    for (auto& x : A) x = (float)(rand() % 7 - 3);
    for (auto& x : B) x = (float)(rand() % 7 - 3);
    std::vector<uint8_t> bi((K / 16) * N * 64, 0);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
        int chunk = k / 16, kk = k % 16, unit = kk / 4, w = kk % 4, us = swizzle_unit(unit, n, 64);
        memcpy(&bi[(size_t)chunk * N * 64 + n * 64 + us * 16 + w * 4], &B[n * K + k], 4);
    }
    float *dA, *dout; uint8_t* dB;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, bi.size()); cudaMalloc(&dout, 128 * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, bi.data(), bi.size(), cudaMemcpyHostToDevice);
    probe<<<1, 128, bi.size() + 2048>>>(dA, dB, dout, K, N, (uint32_t)bi.size());
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> out(128 * N);
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
        double r = 0; for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * B[n * K + k];
        maxerr = fmax(maxerr, fabs(out[m * N + n] - r));
    }
    printf("TS probe (A in TMEM via tcgen05.st, lane=row, col=k): max err %.3f %s\n", maxerr, maxerr == 0 ? "PASS" : "FAIL");
    return 0;
}
