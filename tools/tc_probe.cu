// Probe of the tcgen05 building blocks used by the flow kernels, checked against an fp64 host GEMM:
//   G1  D[128 x N] = A[128 x K] * B[N x K]^T   A from TMEM (tcgen05.st), B from shared memory (K-major),
//       3xTF32 (hi*hi + lo*hi + hi*lo), accumulators in TMEM, read back with tcgen05.ld.
//   G3  D[M x N]   = sum_k A[k][m] * B[k][n]    both operands MN-major in shared memory (weight gradient).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc_probe tools/tc_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#include "../go_with_the_flows_b200/csrc/gwtf_tc.cuh"

using namespace gwtf;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int KP = 40;   // padded K (multiple of 8)
constexpr int NP = 48;   // padded N (multiple of 16)

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_g1(const float* A, const float* B, float* D) {
    __shared__ __align__(128) float sB[2][NP * KP];     // hi / lo, canonical K-major
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    // stage B (N x K row-major in global) into the canonical layout, split hi/lo
    for (int i = tid; i < NP * KP; i += 128) {
        const int n = i / KP, k = i - n * KP;
        float hi, lo;
        split_tf32(B[i], hi, lo);
        const int off = kmajor_offset(n, k, KP);
        sB[0][off] = hi;
        sB[1][off] = lo;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t colD = 0, colAhi = 64, colAlo = 64 + KP;
    // A row of this thread -> TMEM (hi, lo)
    {
        float hi[KP], lo[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) split_tf32(A[tid * KP + k], hi[k], lo[k]);
        tmem_st<KP>(tbase + lane_base + colAhi, hi);
        tmem_st<KP>(tbase + lane_base + colAlo, lo);
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = make_idesc_tf32(128, NP, 0, 0);
        const uint64_t bhi = make_smem_desc_kmajor(sB[0], KP), blo = make_smem_desc_kmajor(sB[1], KP);
        bool acc = false;
        for (int pass = 0; pass < 3; ++pass) {
            const uint32_t a = tbase + (pass == 1 ? colAlo : colAhi);
            const uint64_t b = pass == 2 ? blo : bhi;
            for (int s = 0; s < KP / 8; ++s) {
                mma_tf32_ts(tbase + colD, a + 8 * s, b + (uint64_t)((2 * s * 128) >> 4), idesc, acc);
                acc = true;
            }
        }
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    float d[NP];
    tmem_ld<NP>(tbase + lane_base + colD, d);
    tmem_wait_ld();
#pragma unroll
    for (int n = 0; n < NP; ++n) D[tid * NP + n] = d[n];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
}

// ---------------------------------------------------------------------------------------------
// G3: D[m][n] = sum_k A[k][m] * B[k][n], k = 0..127 (points), m < 128, n < 80; operands written by
// "their" thread k exactly like the backward kernel does (row k = one point).
constexpr int M3 = 128, N3 = 96, K3 = 128;
__global__ void __launch_bounds__(128) k_g3(const float* A, const float* B, float* D, int variant) {
    extern __shared__ __align__(1024) float sm3[];
    float* sA[2] = {sm3, sm3 + M3 * K3};
    float* sBm[2] = {sm3 + 2 * M3 * K3, sm3 + 2 * M3 * K3 + N3 * K3};
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 128);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    const bool a_mn = variant & 1, b_mn = variant & 2, swap = variant & 4;
    for (int m = 0; m < M3; ++m) {
        float hi, lo;
        split_tf32(A[tid * M3 + m], hi, lo);
        const int off = a_mn ? mnmajor_sw32_offset(m, tid, K3) : kmajor_offset(m, tid, K3);
        sA[0][off] = hi; sA[1][off] = lo;
    }
    for (int n = 0; n < N3; ++n) {
        float hi, lo;
        split_tf32(B[tid * N3 + n], hi, lo);
        const int off = b_mn ? mnmajor_sw32_offset(n, tid, K3) : kmajor_offset(n, tid, K3);
        sBm[0][off] = hi; sBm[1][off] = lo;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    const uint32_t tbase = tmem_base_s;
    if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = make_idesc_tf32(M3, N3, a_mn ? 1 : 0, b_mn ? 1 : 0);
        bool acc = false;
        for (int pass = 0; pass < 3; ++pass) {
            const uint32_t kst = 128u, mnst = (uint32_t)(K3 / 8) * 128u;
            const float* pa = sA[pass == 1 ? 1 : 0];
            const float* pb = sBm[pass == 2 ? 1 : 0];
            (void)kst; (void)mnst;
            const uint32_t lb = (uint32_t)(K3 / 4) * 512u, sb = 512u;
            const uint64_t a = !a_mn ? make_smem_desc_kmajor(pa, K3) : (swap ? make_smem_desc(pa, sb, lb, 1u) : make_smem_desc(pa, lb, sb, 1u));
            const uint64_t b = !b_mn ? make_smem_desc_kmajor(pb, K3) : (swap ? make_smem_desc(pb, sb, lb, 1u) : make_smem_desc(pb, lb, sb, 1u));
            for (int s = 0; s < K3 / 8; ++s) {
                const uint64_t ao = a_mn ? (uint64_t)((s * 1024) >> 4) : (uint64_t)((2 * s * 128) >> 4);
                const uint64_t bo = b_mn ? (uint64_t)((s * 1024) >> 4) : (uint64_t)((2 * s * 128) >> 4);
                mma_tf32_ss(tbase, a + ao, b + bo, idesc, acc);
                acc = true;
            }
        }
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    float d[N3];
    tmem_ld<N3>(tbase + ((uint32_t)(warp * 32) << 16), d);
    tmem_wait_ld();
#pragma unroll
    for (int n = 0; n < N3; ++n) D[tid * N3 + n] = d[n];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 128);
}

// ---------------------------------------------------------------------------------------------
// cycle breakdown of one "contraction call" (P tiles): st | sync | mma+commit+wait | ld
__global__ void __launch_bounds__(128) k_time(const float* A, const float* B, float* D, long long* cyc, int iters, int P) {
    __shared__ __align__(128) float sB[2][NP * KP];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    for (int i = tid; i < NP * KP; i += 128) {
        const int n = i / KP, k = i - n * KP;
        float hi, lo;
        split_tf32(B[i], hi, lo);
        const int off = kmajor_offset(n, k, KP);
        sB[0][off] = hi; sB[1][off] = lo;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float hi[KP], lo[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) split_tf32(A[tid * KP + k], hi[k], lo[k]);
    long long t_st = 0, t_sync = 0, t_mma = 0, t_ld = 0;
    uint32_t phase = 0;
    float sink = 0.f;
    for (int it = 0; it < iters; ++it) {
        long long c0 = clock64();
        for (int p = 0; p < P; ++p) {
            tmem_st<KP>(tbase + lane_base + p * 128 + 48, hi);
            tmem_st<KP>(tbase + lane_base + p * 128 + 88, lo);
        }
        tmem_wait_st();
        long long c1 = clock64();
        tc_fence_before();
        __syncthreads();
        long long c2 = clock64();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t idesc = make_idesc_tf32(128, NP, 0, 0);
            const uint64_t bhi = make_smem_desc_kmajor(sB[0], KP), blo = make_smem_desc_kmajor(sB[1], KP);
            for (int p = 0; p < P; ++p) {
                bool acc = false;
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t a = tbase + p * 128 + (pass == 1 ? 88 : 48);
                    const uint64_t b = pass == 2 ? blo : bhi;
                    for (int s = 0; s < KP / 8; ++s) { mma_tf32_ts(tbase + p * 128, a + 8 * s, b + (uint64_t)(16 * s), idesc, acc); acc = true; }
                }
            }
            tc_commit(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1u;
        tc_fence_after();
        long long c3 = clock64();
        for (int p = 0; p < P; ++p) {
            float d[NP];
            tmem_ld<NP>(tbase + lane_base + p * 128, d);
            tmem_wait_ld();
#pragma unroll
            for (int n = 0; n < NP; ++n) sink += d[n];
        }
        long long c4 = clock64();
        t_st += c1 - c0; t_sync += c2 - c1; t_mma += c3 - c2; t_ld += c4 - c3;
    }
    if (tid == 0 && blockIdx.x == 0) { cyc[0] = t_st; cyc[1] = t_sync; cyc[2] = t_mma; cyc[3] = t_ld; }
    D[blockIdx.x * 128 + tid] = sink;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
}

static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }

int main() {
    srand(1);
    {   // ---- G1
        std::vector<float> A(128 * KP), B(NP * KP), D(128 * NP);
        for (auto& v : A) v = (float)frand();
        for (auto& v : B) v = (float)frand();
        float *dA, *dB, *dD;
        CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
        k_g1<<<1, 128>>>(dA, dB, dD);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < NP; ++n) {
                double r = 0;
                for (int k = 0; k < KP; ++k) r += (double)A[m * KP + k] * (double)B[n * KP + k];
                maxerr = fmax(maxerr, fabs(r - D[m * NP + n]));
                maxref = fmax(maxref, fabs(r));
            }
        printf("G1 (TS, K-major B, 3xTF32): max abs err %.3e  (max |ref| %.3f)  %s\n", maxerr, maxref,
               maxerr < 2e-5 ? "OK" : "FAIL");
    }
    {   // ---- G3
        std::vector<float> A(K3 * M3), B(K3 * N3), D(M3 * N3);
        for (auto& v : A) v = (float)frand();
        for (auto& v : B) v = (float)frand();
        float *dA, *dB, *dD;
        CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
        const size_t smem = (size_t)(2 * M3 * K3 + 2 * N3 * K3) * 4;
        CK(cudaFuncSetAttribute(k_g3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int variant = 0; variant < 4; ++variant) {
        CK(cudaMemset(dD, 0xff, D.size() * 4));
        k_g3<<<1, 128, smem>>>(dA, dB, dD, variant);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < M3; ++m)
            for (int n = 0; n < N3; ++n) {
                double r = 0;
                for (int k = 0; k < K3; ++k) r += (double)A[k * M3 + m] * (double)B[k * N3 + n];
                maxerr = fmax(maxerr, fabs(r - D[m * N3 + n]));
                maxref = fmax(maxref, fabs(r));
            }
        printf("G3 variant %d (SS, MN-major A and B, 3xTF32): max abs err %.3e  (max |ref| %.3f)  %s\n", variant, maxerr, maxref,
               maxerr < 5e-5 ? "OK" : "FAIL");
        printf("   D[0][0..3] = %g %g %g %g   D[5][7]=%g\n", D[0], D[1], D[2], D[3], D[5 * N3 + 7]);
        }
    }
    {   // ---- timing
        std::vector<float> A(128 * KP), B(NP * KP);
        for (auto& v : A) v = (float)frand();
        for (auto& v : B) v = (float)frand();
        float *dA, *dB, *dD; long long* dC;
        CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, 148 * 4 * 128 * 4)); CK(cudaMalloc(&dC, 64));
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
        for (int grid : {1, 148, 296}) for (int P : {1, 2}) {
            const int iters = 200;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            k_time<<<grid, 128>>>(dA, dB, dD, dC, iters, P);
            cudaEventRecord(e0);
            k_time<<<grid, 128>>>(dA, dB, dD, dC, iters, P);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long c[4]; CK(cudaMemcpy(c, dC, 32, cudaMemcpyDeviceToHost));
            printf("time grid %3d P %d: %.3f ms total, per call: st %lld sync %lld mma+wait %lld ld %lld cycles (%.2f us/call)\n", grid, P, ms,
                   c[0] / iters, c[1] / iters, c[2] / iters, c[3] / iters, ms * 1e3 / iters);
        }
    }
    return 0;
}
