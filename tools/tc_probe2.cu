// Probe for the tcgen05 backward design (round 2): point-contraction GEMMs D[M x N] = sum_k A[k][m] B[k][n] with
// both operands MN-major in shared memory (written row by row by "their" point thread, 128B-swizzle / 32B-base
// layout), for M = 64 and M = 128 -- where do the rows of an M = 64 accumulator live in tensor memory? -- and the
// rate of such MMAs (shared-memory operand fetch bound?) and of the operand stores.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc_probe2 tools/tc_probe2.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#include "../go_with_the_flows_b200/csrc/gwtf_tc.cuh"

using namespace gwtf;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int K3 = 128;

// D = A^T B, A: [K3][MR] (MR real rows, padded to M), B: [K3][NR]; dumps all 128 lanes x N columns
template <int M, int N>
__global__ void __launch_bounds__(128) k_g(const float* A, const float* B, float* D, int MR, int NR) {
    extern __shared__ __align__(1024) float sm[];
    float* sA[2] = {sm, sm + 128 * K3};
    float* sB[2] = {sm + 2 * 128 * K3, sm + 2 * 128 * K3 + 64 * K3};
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 128);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    for (int i = tid; i < 3 * 128 * K3; i += 128) sm[i] = 0.f;
    __syncthreads();
    for (int m = 0; m < MR; ++m) {
        float hi, lo;
        split_tf32(A[tid * MR + m], hi, lo);
        const int off = mnmajor_sw32_offset(m, tid, K3);
        sA[0][off] = hi; sA[1][off] = lo;
    }
    for (int n = 0; n < NR; ++n) {
        float hi, lo;
        split_tf32(B[tid * NR + n], hi, lo);
        const int off = mnmajor_sw32_offset(n, tid, K3);
        sB[0][off] = hi; sB[1][off] = lo;
    }
    // poison the accumulator so untouched lanes show
    {
        float z[N];
        for (int i = 0; i < N; ++i) z[i] = -777.f;
        tmem_st<N>(tmem_base_s + ((uint32_t)(warp * 32) << 16), z);
        tmem_wait_st();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    const uint32_t tbase = tmem_base_s;
    if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = make_idesc_tf32(M, N, 1, 1);
        bool acc = false;
        for (int pass = 0; pass < 3; ++pass) {
            const uint64_t a = make_smem_desc_mnmajor_sw32(sA[pass == 1 ? 1 : 0], K3);
            const uint64_t b = make_smem_desc_mnmajor_sw32(sB[pass == 2 ? 1 : 0], K3);
            for (int s = 0; s < K3 / 8; ++s) {
                mma_tf32_ss(tbase, a + (uint64_t)((s * 1024) >> 4), b + (uint64_t)((s * 1024) >> 4), idesc, acc);
                acc = true;
            }
        }
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    float d[N];
    tmem_ld<N>(tbase + ((uint32_t)(warp * 32) << 16), d);
    tmem_wait_ld();
    for (int n = 0; n < N; ++n) D[tid * N + n] = d[n];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 128);
}

// timing: NI issuer threads (one per warp 4..4+NI-1) each run `reps` GEMMs of 16 K-steps x 3 passes on their own
// accumulator; operands already in shared memory.  cyc[0] = cycles of the whole loop (issuer 0).
template <int M, int N>
__global__ void __launch_bounds__(256) k_time(long long* cyc, int reps, int NI, int ts_mode) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); mbar_fence_init(); }
    for (int i = tid; i < (128 + 96) * K3; i += 256) sm[i] = 1.0f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    if (warp >= 4 && warp < 4 + NI && (tid & 31) == 0) {
        const int w = warp - 4;
        const uint32_t idesc = ts_mode ? make_idesc_tf32(M, N, 0, 0) : make_idesc_tf32(M, N, 1, 1);
        const uint64_t a = make_smem_desc_mnmajor_sw32(sm, K3);
        const uint64_t b = ts_mode ? make_smem_desc_kmajor(sm + 128 * K3, 40) : make_smem_desc_mnmajor_sw32(sm + 128 * K3, K3);
        uint32_t ph = 0;
        const long long c0 = clock64();
        for (int r = 0; r < reps; ++r) {
            for (int pass = 0; pass < 3; ++pass)
                for (int s = 0; s < (ts_mode ? 5 : 16); ++s) {
                    if (ts_mode) mma_tf32_ts(tbase + w * 128, tbase + w * 128 + 48 + 8 * s, b + (uint64_t)(16 * s), idesc, true);
                    else mma_tf32_ss(tbase + w * 128, a + (uint64_t)((s * 1024) >> 4), b + (uint64_t)((s * 1024) >> 4), idesc, true);
                }
            tc_commit(&bar[w]);
            mbar_wait(&bar[w], ph);
            ph ^= 1u;
        }
        const long long c1 = clock64();
        if (w == 0 && blockIdx.x == 0) cyc[0] = c1 - c0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

// operand-store timing: 128 threads write their 40-float row (hi and lo) into the sw32 MN-major layout
__global__ void __launch_bounds__(128) k_sts(long long* cyc, float* sink, int reps, int swap_halves) {
    extern __shared__ __align__(1024) float sm[];
    const int tid = threadIdx.x;
    float v[40];
    for (int i = 0; i < 40; ++i) v[i] = tid * 0.01f + i;
    __syncthreads();
    const long long c0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int arr = 0; arr < 2; ++arr) {
            float* base = sm + arr * 8192;
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                const int off = mnmajor_sw32_offset(8 * c, tid, K3);       // 32-byte chunk of channels 8c..8c+7
                const bool sw = swap_halves && ((tid >> 2) & 1);
                const float4 lo4 = make_float4(v[8 * c] + r, v[8 * c + 1], v[8 * c + 2], v[8 * c + 3]);
                const float4 hi4 = make_float4(v[8 * c + 4] + r, v[8 * c + 5], v[8 * c + 6], v[8 * c + 7]);
                if (swap_halves) {
                    // half of the lanes write their upper 16 bytes first: 8 distinct bank groups per instruction
                    *reinterpret_cast<float4*>(base + off + (sw ? 4 : 0)) = sw ? hi4 : lo4;
                    *reinterpret_cast<float4*>(base + off + (sw ? 0 : 4)) = sw ? lo4 : hi4;
                } else {
                    *reinterpret_cast<float4*>(base + off) = lo4;
                    *reinterpret_cast<float4*>(base + off + 4) = hi4;
                }
            }
        }
        __syncthreads();
    }
    const long long c1 = clock64();
    if (tid == 0) cyc[0] = c1 - c0;
    sink[tid] = sm[tid];
}

static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }

template <int M, int N>
int run_g(int MR, int NR) {
    std::vector<float> A(K3 * MR), B(K3 * NR), D(128 * N);
    for (auto& v : A) v = (float)frand();
    for (auto& v : B) v = (float)frand();
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    const size_t smem = (size_t)3 * 128 * K3 * 4;
    CK(cudaFuncSetAttribute(k_g<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_g<M, N><<<1, 128, smem>>>(dA, dB, dD, MR, NR);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    // which lane holds row m?  search
    std::vector<int> lane_of(MR, -1);
    double maxerr = 0;
    for (int m = 0; m < MR; ++m) {
        std::vector<double> ref(NR);
        for (int n = 0; n < NR; ++n) {
            double r = 0;
            for (int k = 0; k < K3; ++k) r += (double)A[k * MR + m] * (double)B[k * NR + n];
            ref[n] = r;
        }
        double best = 1e30; int bl = -1;
        for (int l = 0; l < 128; ++l) {
            double e = 0;
            for (int n = 0; n < NR; ++n) e = fmax(e, fabs(ref[n] - D[l * N + n]));
            if (e < best) { best = e; bl = l; }
        }
        lane_of[m] = bl;
        maxerr = fmax(maxerr, best);
    }
    printf("G M=%d N=%d (MR=%d NR=%d): max abs err %.3e %s; row->lane:", M, N, MR, NR, maxerr, maxerr < 5e-5 ? "OK" : "FAIL");
    for (int m = 0; m < MR; m += (MR > 16 ? 8 : 1)) printf(" %d->%d", m, lane_of[m]);
    int untouched = 0;
    for (int l = 0; l < 128; ++l) if (D[l * N] == -777.f) ++untouched;
    printf("  (lanes untouched: %d)\n", untouched);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return 0;
}

template <int M, int N>
int run_time(int ts_mode) {
    long long* dC;
    CK(cudaMalloc(&dC, 64));
    const size_t smem = (size_t)(128 + 96) * K3 * 4;
    CK(cudaFuncSetAttribute(k_time<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int NI : {1, 2, 4}) {
        const int reps = 200;
        k_time<M, N><<<148, 256, smem>>>(dC, reps, NI, ts_mode);
        CK(cudaDeviceSynchronize());
        long long c;
        CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
        const int n_mma = ts_mode ? 15 : 48;
        printf("time %s M=%d N=%d issuers=%d: %.0f cycles per GEMM per issuer (%d MMAs) -> %.1f cycles/MMA aggregate\n",
               ts_mode ? "TS(K=40)" : "SS(K=128pts)", M, N, NI, (double)c / reps, n_mma, (double)c / reps / n_mma / NI);
    }
    cudaFree(dC);
    return 0;
}

int main() {
    srand(2);
    if (run_g<128, 48>(40, 40)) return 1;
    if (run_g<64, 48>(40, 40)) return 1;
    if (run_g<64, 40>(40, 40)) return 1;
    if (run_g<64, 8>(64, 8)) return 1;
    if (run_g<128, 16>(120, 8)) return 1;
    if (run_time<64, 40>(0)) return 1;
    if (run_time<128, 48>(0)) return 1;
    if (run_time<64, 8>(0)) return 1;
    if (run_time<128, 16>(0)) return 1;
    if (run_time<64, 80>(0)) return 1;
    if (run_time<128, 48>(1)) return 1;
    if (run_time<128, 16>(1)) return 1;
    {
        long long* dC; float* dS;
        CK(cudaMalloc(&dC, 64)); CK(cudaMalloc(&dS, 1024));
        const size_t smem = (size_t)2 * 128 * K3 * 4;
        CK(cudaFuncSetAttribute(k_sts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int sw : {0, 1}) {
            k_sts<<<1, 128, smem>>>(dC, dS, 200, sw);
            CK(cudaDeviceSynchronize());
            long long c;
            CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
            printf("operand stores (128 threads x 2 arrays x 40 floats, swap_halves=%d): %.0f cycles per tile\n", sw, (double)c / 200);
        }
    }
    return 0;
}
