// Scan which smem byte offset feeds A[m][k] for MN-major A descriptors (one-hot flooding).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../dl4ds_b200/csrc/tc_common.cuh"
using namespace dl4ds::tc;
namespace dl4ds { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

struct Probe { uint32_t a_lbo, a_sbo, a_layout, idesc; uint32_t hot; float val; uint32_t region; };

__global__ void scan_kernel(const uint8_t* b_img, float* out, Probe p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    for (uint32_t i = threadIdx.x; i < p.region / 4; i += blockDim.x) ((float*)sm)[i] = (p.hot == 0xffffffffu) ? p.val : (i == p.hot ? p.val : 0.f);
    uint8_t* smb = sm + p.region;
    for (uint32_t i = threadIdx.x; i < 512 / 4; i += blockDim.x) ((uint32_t*)smb)[i] = ((const uint32_t*)b_img)[i];
    fence_proxy_async_smem();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 32);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t td = tmem_slot;
    if (threadIdx.x == 0) {
        uint64_t da = make_smem_desc(base, p.a_lbo, p.a_sbo, p.a_layout);
        uint64_t db = make_smem_desc(base + p.region, 16, 256, kLayoutSw32);   // K-major SW32 B: 16 rows x 32B
        umma_tf32(td, da, db, p.idesc, 0);
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    const uint32_t taddr = td + ((uint32_t)(warp * 32) << 16);
    float v[16];
    tmem_ld16(taddr, v);
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = v[j];
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(td, 32);
}

int main(int argc, char** argv) {
    uint8_t* db; float* dout;
    cudaMalloc(&db, 4096); cudaMalloc(&dout, 128 * 16 * 4);
    cudaFuncSetAttribute(scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    // B (N=16 x K=8) K-major SW32: B[n][k] = (k+1) for n==0, else 0; so D[m][0] = sum_k A[m][k]*(k+1) -> identifies k
    std::vector<uint8_t> bi(512, 0);
    for (int k = 0; k < 8; ++k) { float v = (float)(k + 1); int unit = k / 4, w = k % 4; int us = swizzle_unit(unit, 0, 32); memcpy(&bi[us * 16 + w * 4], &v, 4); }
    cudaMemcpy(db, bi.data(), 512, cudaMemcpyHostToDevice);
    const uint32_t region = 32768;
    std::vector<float> out(128 * 16);
    for (int amn = 0; amn < 2; ++amn)
    for (uint32_t layout : {kLayoutSw32, kLayoutSw64, kLayoutSw128, 0u}) {
        Probe p; p.a_lbo = 4096; p.a_sbo = 2048; p.a_layout = layout; p.idesc = make_idesc_tf32(64, 16, amn, 0); p.region = region; p.val = 1.0f;
        // flood
        p.hot = 0xffffffffu;
        scan_kernel<<<1, 128, region + 512 + 2048>>>(db, dout, p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        printf("a_major %d layout %u flood: D[0][0]=%g D[1][0]=%g D[17(lane 33)][0]=%g (expect 36 if all in range)\n", amn, layout, out[0], out[16], out[33 * 16]);
        // one-hot scan
        std::vector<int> off(64 * 8, -1);
        int found = 0;
        for (uint32_t h = 0; h < region / 4; ++h) {
            p.hot = h;
            scan_kernel<<<1, 128, region + 512 + 2048>>>(db, dout, p);
            cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
            for (int m = 0; m < 64; ++m) {
                int lane = 32 * (m / 16) + m % 16;
                float g = out[lane * 16];
                if (g != 0) { int k = (int)g - 1; if (k >= 0 && k < 8) { off[m * 8 + k] = (int)h * 4; found++; } }
            }
        }
        printf("  found %d of 512 elements. byte offset of A[m][k]:\n", found);
        for (int m : {0, 1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63}) {
            printf("   m=%2d:", m);
            for (int k = 0; k < 8; ++k) printf(" %6d", off[m * 8 + k]);
            printf("\n");
        }
    }
    return 0;
}
