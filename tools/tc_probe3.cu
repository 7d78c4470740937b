// Is the ~120-cycle cost per small tcgen05.mma an issue overhead of the issuing thread, or the latency of a chain of
// MMAs that accumulate into the SAME tensor-memory tile?  One issuer thread, NACC accumulators used round-robin.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../go_with_the_flows_b200/csrc/gwtf_tc.cuh"
using namespace gwtf;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int N>
__global__ void __launch_bounds__(128) k_rr(long long* cyc, int reps, int nacc, int ts_mode, int n_mma) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    for (int i = tid; i < 48 * 1024 / 4; i += 128) sm[i] = 1.0f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    if (tid == 32) {
        const uint32_t idesc = make_idesc_tf32(128, N, 0, 0);
        const uint64_t a = make_smem_desc_kmajor(sm, 40);
        const uint64_t b = make_smem_desc_kmajor(sm + 128 * 40, 40);
        uint32_t ph = 0;
        const long long c0 = clock64();
        for (int r = 0; r < reps; ++r) {
            for (int i = 0; i < n_mma; ++i) {
                const int acc = i % nacc, s = (i / nacc) % 5;
                const uint32_t d = tbase + acc * 64;                 // accumulators 64 columns apart (N <= 48)
                if (ts_mode) mma_tf32_ts(d, tbase + 448 + 8 * s, b + (uint64_t)(16 * s), idesc, true);
                else mma_tf32_ss(d, a + (uint64_t)(16 * s), b + (uint64_t)(16 * s), idesc, true);
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

int main() {
    long long* dC;
    CK(cudaMalloc(&dC, 64));
    const size_t smem = 48 * 1024;
    CK(cudaFuncSetAttribute(k_rr<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_rr<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int ts : {1, 0})
        for (int n_mma : {15, 60})
            for (int nacc : {1, 2, 3, 4, 6}) {
                const int reps = 200;
                k_rr<48><<<148, 128, smem>>>(dC, reps, nacc, ts, n_mma);
                CK(cudaDeviceSynchronize());
                long long c; CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
                k_rr<16><<<148, 128, smem>>>(dC, reps, nacc, ts, n_mma);
                CK(cudaDeviceSynchronize());
                long long c16; CK(cudaMemcpy(&c16, dC, 8, cudaMemcpyDeviceToHost));
                printf("%s batch of %2d MMAs over %d accumulators: N=48 %.0f cycles/batch (%.1f per MMA)   N=16 %.0f (%.1f per MMA)\n",
                       ts ? "TS" : "SS", n_mma, nacc, (double)c / reps, (double)c / reps / n_mma, (double)c16 / reps, (double)c16 / reps / n_mma);
            }
    return 0;
}
