// Throughput of the warp-level tensor-core path on this GPU: mma.sync.m16n8k8 tf32 (register fragments).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_peak tools/mma_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int NACC>
__global__ void __launch_bounds__(256) k_mma(float* out, int iters) {
    float d[NACC][4];
    uint32_t a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
    for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + i);
#pragma unroll
    for (int n = 0; n < NACC; ++n) for (int i = 0; i < 4; ++i) d[n][i] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int n = 0; n < NACC; ++n)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[n][0]), "+f"(d[n][1]), "+f"(d[n][2]), "+f"(d[n][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
#pragma unroll
    for (int n = 0; n < NACC; ++n) for (int i = 0; i < 4; ++i) s += d[n][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 4000;
    for (int bps : {1, 2, 4}) {
        int grid = 148 * bps;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_mma<16><<<grid, 256>>>(out, iters);
        cudaEventRecord(e0);
        k_mma<16><<<grid, 256>>>(out, iters);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 16 * 8 * 8 * 16.0 * iters * 8.0 * grid;   // per warp-instruction 16x8x8 MAC, 8 warps/CTA
        printf("mma.sync m16n8k8 tf32, %d CTA/SM x 8 warps: %.3f ms, %.1f TFLOP/s (3xTF32 effective %.1f)\n", bps, ms, flops / ms * 1e-9, flops / ms * 1e-9 / 3);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
