// Issue rate of the legacy warp-level tensor-core path on sm_100a: mma.sync m16n8k8 tf32 vs m16n8k16 f16 / bf16.
// Each warp runs `iters` rounds of NACC independent accumulator chains; clocks per MMA per SM sub-partition are
// reported for 1..8 warps per sub-partition.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_rate mma_sync_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND, int NACC>
__global__ void rate_kernel(float* out, long long* clk, int iters) {
    float acc[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 5, b1 = 11;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int KIND, int NACC>
void run(const char* name, float* out, long long* clk) {
    const int iters = 4096;
    for (int warps = 4; warps <= 32; warps *= 2) {
        rate_kernel<KIND, NACC><<<148, warps * 32>>>(out, clk, iters);
        cudaDeviceSynchronize();
        rate_kernel<KIND, NACC><<<148, warps * 32>>>(out, clk, iters);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        const double per_smsp = (double)mx / ((double)iters * NACC * (warps / 4));
        printf("%-22s acc-chains %d  warps/SM %2d : %7.2f clk per MMA per sub-partition (latency-bound if chains*warps small)\n",
               name, NACC, warps, per_smsp);
    }
}

int main() {
    float* out; long long* clk;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
    run<0, 1>("m16n8k8 tf32", out, clk);  run<0, 4>("m16n8k8 tf32", out, clk);
    run<1, 1>("m16n8k16 f16", out, clk);  run<1, 4>("m16n8k16 f16", out, clk);
    run<2, 4>("m16n8k16 bf16", out, clk);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
