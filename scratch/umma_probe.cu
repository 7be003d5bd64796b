// UMMA layout probe: one CTA, operands written to smem by threads from host-built images.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../dl4ds_b200/csrc/tc_common.cuh"
using namespace dl4ds::tc;
namespace dl4ds { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

struct Probe { uint32_t a_bytes, b_bytes; uint32_t a_lbo, a_sbo, a_layout, b_lbo, b_sbo, b_layout; uint32_t idesc; int ncols; int nk; uint32_t a_kstep, b_kstep; };

__global__ void probe_kernel(const uint8_t* a_img, const uint8_t* b_img, float* out, Probe p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    for (uint32_t i = threadIdx.x; i < p.a_bytes / 4; i += blockDim.x) ((uint32_t*)sm)[i] = ((const uint32_t*)a_img)[i];
    uint8_t* smb = sm + ((p.a_bytes + 1023) & ~1023u);
    for (uint32_t i = threadIdx.x; i < p.b_bytes / 4; i += blockDim.x) ((uint32_t*)smb)[i] = ((const uint32_t*)b_img)[i];
    fence_proxy_async_smem();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 256);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t td = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t sa = base, sb = base + ((p.a_bytes + 1023) & ~1023u);
        for (int k = 0; k < p.nk; ++k) {
            uint64_t da = make_smem_desc(sa + k * p.a_kstep, p.a_lbo, p.a_sbo, p.a_layout);
            uint64_t db = make_smem_desc(sb + k * p.b_kstep, p.b_lbo, p.b_sbo, p.b_layout);
            umma_tf32(td, da, db, p.idesc, k > 0);
        }
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    const uint32_t taddr = td + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < p.ncols; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);
        for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * p.ncols + c0 + j] = v[j];
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(td, 256);
}

static int swz_unit(int unit, int row, int span) { return swizzle_unit(unit, row, span); }

// K-major image: rows = MN index, span bytes (kc = span/4 k-elements) per row; tile of K = nk*8 split in chunks of kc? here K <= kc.
static void img_kmajor(std::vector<uint8_t>& img, const std::vector<float>& mat, int MN, int K, int span) {
    img.assign((size_t)MN * span, 0);
    for (int r = 0; r < MN; ++r) for (int k = 0; k < K; ++k) {
        int unit = k / 4, w = k % 4;
        int us = swz_unit(unit, r, span);
        memcpy(&img[(size_t)r * span + us * 16 + w * 4], &mat[(size_t)r * K + k], 4);
    }
}
// MN-major image: rows = K index (pixels), each row `span` bytes = kc MN-elements; MN blocks of kc at block_stride bytes
static void img_mnmajor(std::vector<uint8_t>& img, const std::vector<float>& mat, int MN, int K, int span, int block_stride, int nblocks_alloc) {
    int kc = span / 4;
    img.assign((size_t)nblocks_alloc * block_stride, 0);
    for (int m = 0; m < MN; ++m) for (int k = 0; k < K; ++k) {
        int blk = m / kc, mi = m % kc, unit = mi / 4, w = mi % 4;
        int us = swz_unit(unit, k, span);
        memcpy(&img[(size_t)blk * block_stride + (size_t)k * span + us * 16 + w * 4], &mat[(size_t)m * K + k], 4);
    }
}
static uint32_t layout_of(int span) { return span == 128 ? kLayoutSw128 : span == 64 ? kLayoutSw64 : kLayoutSw32; }

int main() {
    uint8_t *da, *db; float* dout;
    cudaMalloc(&da, 1 << 20); cudaMalloc(&db, 1 << 20); cudaMalloc(&dout, 128 * 256 * 4);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    srand(1);
    struct Case { const char* name; int M, N, K; int a_mn, b_mn; int a_span, b_span; int variant; };
    std::vector<Case> cases;
    for (int span : {32, 64, 128}) {
        cases.push_back({"Kmaj/Kmaj M128", 128, 16, 8, 0, 0, span, span, 0});
        cases.push_back({"Kmaj/Kmaj M64", 64, 16, 8, 0, 0, span, span, 0});
        for (int variant = 0; variant < 2; ++variant) {
            cases.push_back({"MNmaj/MNmaj M64 K8", 64, 16, 8, 1, 1, span, span, variant});
            cases.push_back({"MNmaj/MNmaj M64 K64", 64, 96, 64, 1, 1, span, span, variant});
            cases.push_back({"MNmaj/MNmaj M128 K16", 128, 32, 16, 1, 1, span, span, variant});
            cases.push_back({"MNmaj A / Kmaj B M64", 64, 16, 8, 1, 0, span, span, variant});
            cases.push_back({"Kmaj A / MNmaj B M128", 128, 48, 8, 0, 1, span, span, variant});
        }
    }
    for (auto& c : cases) {
        std::vector<float> A((size_t)c.M * c.K), B((size_t)c.N * c.K);
        for (auto& x : A) x = (float)(rand() % 7 - 3);
        for (auto& x : B) x = (float)(rand() % 7 - 3);
        std::vector<uint8_t> ai, bi;
        Probe p; memset(&p, 0, sizeof(p));
        p.nk = c.K / 8;
        const int KT = c.K;        // rows (pixels) per MN-major box
        if (c.a_mn) {
            int bs = KT * c.a_span; if (bs < 1024) bs = 1024;
            int nb = (c.M * 4 + c.a_span - 1) / c.a_span;
            img_mnmajor(ai, A, c.M, c.K, c.a_span, bs, nb);
            p.a_lbo = c.variant == 0 ? bs : 8 * c.a_span; p.a_sbo = c.variant == 0 ? 8 * c.a_span : bs;
            p.a_kstep = 8 * c.a_span;
        } else {
            if (c.K * 4 > c.a_span) continue;
            img_kmajor(ai, A, c.M, c.K, c.a_span);
            p.a_lbo = 16; p.a_sbo = 8 * c.a_span; p.a_kstep = 32;
        }
        if (c.b_mn) {
            int bs = KT * c.b_span; if (bs < 1024) bs = 1024;
            int nb = (c.N * 4 + c.b_span - 1) / c.b_span;
            img_mnmajor(bi, B, c.N, c.K, c.b_span, bs, nb);
            p.b_lbo = c.variant == 0 ? bs : 8 * c.b_span; p.b_sbo = c.variant == 0 ? 8 * c.b_span : bs;
            p.b_kstep = 8 * c.b_span;
        } else {
            if (c.K * 4 > c.b_span) continue;
            img_kmajor(bi, B, c.N, c.K, c.b_span);
            p.b_lbo = 16; p.b_sbo = 8 * c.b_span; p.b_kstep = 32;
        }
        p.a_layout = layout_of(c.a_span); p.b_layout = layout_of(c.b_span);
        p.a_bytes = (uint32_t)((ai.size() + 15) & ~15u); p.b_bytes = (uint32_t)((bi.size() + 15) & ~15u);
        ai.resize(p.a_bytes); bi.resize(p.b_bytes);
        p.idesc = make_idesc_tf32(c.M, c.N, c.a_mn, c.b_mn);
        p.ncols = (c.N + 15) & ~15;
        cudaMemcpy(da, ai.data(), p.a_bytes, cudaMemcpyHostToDevice);
        cudaMemcpy(db, bi.data(), p.b_bytes, cudaMemcpyHostToDevice);
        cudaMemset(dout, 0xff, 128 * 256 * 4);
        size_t smem = ((p.a_bytes + 1023) & ~1023u) + p.b_bytes + 2048 + 32768;
        probe_kernel<<<1, 128, smem>>>(da, db, dout, p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-28s span %3d var %d: CUDA error %s\n", c.name, c.a_span, c.variant, cudaGetErrorString(e)); return 1; }
        std::vector<float> out(128 * p.ncols);
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0; int nz = 0;
        for (int m = 0; m < c.M; ++m) for (int n = 0; n < c.N; ++n) {
            double r = 0; for (int k = 0; k < c.K; ++k) r += (double)A[(size_t)m * c.K + k] * B[(size_t)n * c.K + k];
            int lane = c.M == 128 ? m : 32 * (m / 16) + m % 16;
            double g = out[(size_t)lane * p.ncols + n];
            if (g != 0) nz++;
            double d = fabs(g - r); if (d > maxerr) maxerr = d;
        }
        printf("%-28s span %3d var %d (lbo %5u sbo %5u): max err %8.2f nonzero %d %s\n", c.name, c.a_span, c.variant, c.a_mn ? p.a_lbo : p.b_lbo, c.a_mn ? p.a_sbo : p.b_sbo, maxerr, nz, maxerr == 0 ? "PASS" : "FAIL");
    }
    return 0;
}
