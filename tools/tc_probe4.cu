// What does ONE tcgen05.mma cost the issuing thread?  Variants of the issue loop:
//   0: 16 identical MMAs (no address arithmetic)        1: descriptors advanced by compile-time constants (unrolled)
//   2: like 1 but without the per-MMA predicate setp     3: two warps' lane 0 interleaved (reference for scaling)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../go_with_the_flows_b200/csrc/gwtf_tc.cuh"
using namespace gwtf;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void mma_ts_acc(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc) {
    asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 1;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc) : "memory");
}

template <int VAR>
__global__ void __launch_bounds__(128) k_issue(long long* cyc, int reps) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    for (int i = tid; i < 32 * 1024 / 4; i += 128) sm[i] = 1.0f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    if (warp == 1 && elect_one()) {
        const uint32_t idesc = make_idesc_tf32(128, 48, 0, 0);
        const uint64_t b = make_smem_desc_kmajor(sm, 40);
        uint32_t ph = 0;
        const long long c0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int i = 0; i < 15; ++i) {
                const int s = i % 5;
                if (VAR == 0) mma_tf32_ts(tbase, tbase + 448, b, idesc, true);
                else if (VAR == 1) mma_tf32_ts(tbase, tbase + 448 + 8 * s, b + (uint64_t)(16 * s), idesc, true);
                else mma_ts_acc(tbase, tbase + 448 + 8 * s, b + (uint64_t)(16 * s), idesc);
            }
            tc_commit(&bar);
            mbar_wait(&bar, ph);
            ph ^= 1u;
        }
        const long long c1 = clock64();
        if (blockIdx.x == 0) cyc[0] = c1 - c0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

// no wait between batches: issue `reps` batches back to back, one commit at the end
template <int VAR>
__global__ void __launch_bounds__(128) k_stream(long long* cyc, int reps) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    for (int i = tid; i < 32 * 1024 / 4; i += 128) sm[i] = 1.0f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    if (warp == 1 && elect_one()) {
        const uint32_t idesc = make_idesc_tf32(128, VAR, 0, 0);
        const uint64_t b = make_smem_desc_kmajor(sm, 40);
        const long long c0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int i = 0; i < 15; ++i) {
                const int s = i % 5;
                mma_tf32_ts(tbase, tbase + 448 + 8 * s, b + (uint64_t)(16 * s), idesc, true);
            }
        }
        const long long c1 = clock64();
        tc_commit(&bar);
        mbar_wait(&bar, 0);
        const long long c2 = clock64();
        if (blockIdx.x == 0) { cyc[0] = c1 - c0; cyc[1] = c2 - c0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
    long long* dC;
    CK(cudaMalloc(&dC, 64));
    const size_t smem = 32 * 1024;
    const int reps = 200;
    long long c[2];
#define RUN(K, NAME) CK(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    K<<<148, 128, smem>>>(dC, reps); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost)); \
    printf("%-56s %.1f cycles per MMA (issue)  %.1f (incl. drain)\n", NAME, (double)c[0] / reps / 15, (double)c[1] / reps / 15);
    c[1] = 0;
    RUN(k_issue<0>, "batch of 15 + commit + wait, identical operands");
    RUN(k_issue<1>, "batch of 15 + commit + wait, constant-advanced descs");
    RUN(k_issue<2>, "batch of 15 + commit + wait, constant predicate");
    RUN(k_stream<48>, "stream of 3000 MMAs N=48, no waits");
    RUN(k_stream<16>, "stream of 3000 MMAs N=16, no waits");
    RUN(k_stream<96>, "stream of 3000 MMAs N=96, no waits");
    RUN(k_stream<192>, "stream of 3000 MMAs N=192, no waits");
    return 0;
}
