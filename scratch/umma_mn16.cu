// Probe: tcgen05.mma kind::f16 with MN-MAJOR A and B tiles (rows = k, 128-byte rows of 64 halfs, SWIZZLE_128B), the layout
// an NHWC activation tile has when pixels are the GEMM K dimension (weight gradient).  Checks D = A^T-style product
// against the host for a plain start address, a second K-step (start + 16 rows) and a ROW-SHIFTED start (the kernel-tap
// trick of the forward kernel), for a few (LBO, SBO) candidates.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../dl4ds_b200/csrc/tc_common.cuh"
using namespace dl4ds::tc;
namespace dl4ds { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

struct Probe { uint32_t a_lbo, a_sbo, b_lbo, b_sbo, shift_rows, ksteps, N; };

// smem image: A region (rows x 2 panels x 128 B), B region (rows x 128 B); both written by the host in their final
// (swizzled) form
__global__ void probe_kernel(const uint8_t* a_img, const uint8_t* b_img, uint32_t a_bytes, uint32_t b_bytes, float* out, Probe p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    for (uint32_t i = threadIdx.x; i < a_bytes / 4; i += blockDim.x) ((uint32_t*)sm)[i] = ((const uint32_t*)a_img)[i];
    for (uint32_t i = threadIdx.x; i < b_bytes / 4; i += blockDim.x) ((uint32_t*)(sm + a_bytes))[i] = ((const uint32_t*)b_img)[i];
    fence_proxy_async_smem();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 64);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t td = tmem_slot;
    if (threadIdx.x == 0) {
        uint32_t idesc = make_idesc_f16(128, (int)p.N) | (1u << 15) | (1u << 16);     // a_major = b_major = MN
        for (uint32_t k = 0; k < p.ksteps; ++k) {
            uint64_t da = make_smem_desc(base + (p.shift_rows + 16 * k) * 128, p.a_lbo, p.a_sbo, kLayoutSw128);
            uint64_t db = make_smem_desc(base + a_bytes + 16 * k * 128, p.b_lbo, p.b_sbo, kLayoutSw128);
            umma_f16(td, da, db, idesc, k > 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    const uint32_t taddr = td + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 64; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);
        for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = v[j];
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(td, 64);
}

int main() {
    const int ROWS = 48, K = 32, M = 128, N = 48;       // rows in smem (room for shifts), K used per test
    // logical A[row][m], B[row][n] (row = k index in smem)
    std::vector<float> A(ROWS * M), B(ROWS * 64);
    srand(1);
    for (auto& v : A) v = (float)((rand() % 7) - 3);
    for (auto& v : B) v = (float)((rand() % 5) - 2);
    const uint32_t panel = ROWS * 128;                   // bytes of one 64-wide MN panel (all rows)
    std::vector<uint8_t> ai(2 * panel, 0), bi(panel, 0);
    auto put = [](std::vector<uint8_t>& img, uint32_t panel_off, int row, int col, float v) {   // col in [0,64)
        const int unit = col / 8, w = col % 8;
        const uint32_t off = panel_off + row * 128 + ((unit ^ (row & 7)) << 4) + w * 2;
        __half h = __float2half(v);
        memcpy(&img[off], &h, 2);
    };
    for (int r = 0; r < ROWS; ++r) {
        for (int m = 0; m < M; ++m) put(ai, (m / 64) * panel, r, m % 64, A[r * M + m]);
        for (int n = 0; n < 64; ++n) put(bi, 0, r, n, B[r * 64 + n]);
    }
    uint8_t *da, *db; float* dout;
    cudaMalloc(&da, ai.size()); cudaMalloc(&db, bi.size()); cudaMalloc(&dout, 128 * 64 * 4);
    cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    std::vector<float> out(128 * 64);
    const uint32_t cand_lbo[] = {panel, 1024, 128}, cand_sbo[] = {1024, panel, 128};
    for (uint32_t lbo : cand_lbo) for (uint32_t sbo : cand_sbo) for (uint32_t shift : {0u, 3u, 11u}) for (uint32_t ks : {1u, 2u}) {
        Probe p{lbo, sbo, lbo, sbo, shift, ks, (uint32_t)N};
        probe_kernel<<<1, 128, ai.size() + bi.size() + 2048>>>(da, db, (uint32_t)ai.size(), (uint32_t)bi.size(), dout, p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0; int bad = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (uint32_t k = 0; k < 16 * ks; ++k) ref += (double)A[(shift + k) * M + m] * B[k * 64 + n];
            double d = fabs(out[m * 64 + n] - ref);
            if (d > worst) worst = d;
            if (d > 1e-3) ++bad;
        }
        printf("LBO %5u SBO %5u shift %2u ksteps %u: max err %.3g, %d / %d wrong%s\n", lbo, sbo, shift, ks, worst, bad, M * N, bad == 0 ? "   <== OK" : "");
    }
    return 0;
}
