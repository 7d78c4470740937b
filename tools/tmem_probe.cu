// Cost of using tensor memory as per-thread accumulator storage: each warp loads NC columns of its lane
// quadrant (tcgen05.ld 32x32b), adds, stores them back (tcgen05.st), in a loop.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I go_with_the_flows_b200/csrc -o tools/tmem_probe tools/tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "gwtf_common.cuh"
#include "gwtf_tc.cuh"
using namespace gwtf;

template <int NC>
__global__ void __launch_bounds__(512, 1) k_probe(int active_warps, int iters, long long* cycles, float* sink) {
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&tbase, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
    float v[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = 0.f;
    if (warp < active_warps) { tmem_st<NC>(taddr, v); tmem_wait_st(); }
    __syncthreads();
    long long t0 = clock64();
    if (warp < active_warps) {
        for (int it = 0; it < iters; ++it) {
            tmem_ld<NC>(taddr, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < NC; ++i) v[i] += 1.0f;
            tmem_st<NC>(taddr, v);
            tmem_wait_st();
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (warp < active_warps) { float s = 0.f; for (int i = 0; i < NC; ++i) s += v[i]; sink[blockIdx.x * 512 + threadIdx.x] = s; }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
    long long* cyc; float* sink;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
    const int iters = 2000;
    for (int aw : {1, 4, 8, 16}) {
        k_probe<64><<<148, 512>>>(aw, iters, cyc, sink);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("x64 ld+st round trip, %2d warps/SM: %.1f cycles per iteration per warp (%.1f B/cycle/SM each way)\n", aw,
               (double)h / iters, aw * 64.0 * 32 * 4 / ((double)h / iters));
    }
    for (int aw : {4, 16}) {
        k_probe<128><<<148, 512>>>(aw, iters, cyc, sink);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("x128 ld+st round trip, %2d warps/SM: %.1f cycles per iteration per warp (%.1f B/cycle/SM each way)\n", aw,
               (double)h / iters, aw * 128.0 * 32 * 4 / ((double)h / iters));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
