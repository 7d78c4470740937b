// Probe: can the four MN-major (sw32) operand arrays of the dW1 point contraction share ONE second mn-atom?
// Each array keeps channels 0..31 in its own 16 KB atom; channels 32..39 of array a live in 32-byte slot a of every
// 128-byte row of a shared 16 KB atom (descriptor: LBO = shared + 32 a - base_a).  D[64 x 40] = A^T B, 3xTF32.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "../go_with_the_flows_b200/csrc/gwtf_tc.cuh"
using namespace gwtf;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ int shared_atom_offset(int arr, int i, int k) {       // channel 32 + i of array `arr`, point k
    return (k >> 2) * 128 + (k & 3) * 32 + ((arr ^ (k & 3)) << 3) + i;
}

__global__ void __launch_bounds__(128) k_g(const float* A, const float* B, float* D) {
    extern __shared__ __align__(1024) float sm[];
    // [A_hi | A_lo | B_hi | B_lo] atom 0 each (4096 floats), then the shared atom 1 (4096 floats) + slack
    float* arr[4] = {sm, sm + 4096, sm + 8192, sm + 12288};
    float* shared = sm + 16384;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 64);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    for (int i = tid; i < 5 * 4096 + 256; i += 128) sm[i] = 12345.f;       // junk everywhere: only written cells may matter
    __syncthreads();
    for (int m = 0; m < 40; ++m) {
        float ah, al, bh, bl;
        split_tf32(A[tid * 40 + m], ah, al);
        split_tf32(B[tid * 40 + m], bh, bl);
        if (m < 32) {
            const int off = mnmajor_sw32_offset(m, tid, 128);
            arr[0][off] = ah; arr[1][off] = al; arr[2][off] = bh; arr[3][off] = bl;
        } else {
            shared[shared_atom_offset(0, m - 32, tid)] = ah;
            shared[shared_atom_offset(1, m - 32, tid)] = al;
            shared[shared_atom_offset(2, m - 32, tid)] = bh;
            shared[shared_atom_offset(3, m - 32, tid)] = bl;
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    const uint32_t tbase = tmem_base_s;
    if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = make_idesc_tf32(64, 40, 1, 1);
        uint64_t d[4];
        for (int a = 0; a < 4; ++a) {
            const uint32_t lbo = smem_u32(shared + 8 * a) - smem_u32(arr[a]);
            d[a] = make_smem_desc(arr[a], lbo, 512u, 1u);
        }
        bool acc = false;
        for (int pass = 0; pass < 3; ++pass) {
            const uint64_t a = d[pass == 1 ? 1 : 0], b = d[pass == 2 ? 3 : 2];
            for (int s = 0; s < 16; ++s) {
                mma_tf32_ss(tbase, a + (uint64_t)((s * 1024) >> 4), b + (uint64_t)((s * 1024) >> 4), idesc, acc);
                acc = true;
            }
        }
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    float dd[40];
    tmem_ld<40>(tbase + ((uint32_t)(warp * 32) << 16), dd);
    tmem_wait_ld();
    for (int n = 0; n < 40; ++n) D[tid * 40 + n] = dd[n];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 64);
}

static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }

int main() {
    srand(3);
    std::vector<float> A(128 * 40), B(128 * 40), D(128 * 40);
    for (auto& v : A) v = (float)frand();
    for (auto& v : B) v = (float)frand();
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    const size_t smem = (size_t)(5 * 4096 + 256) * 4;
    CK(cudaFuncSetAttribute(k_g, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_g<<<1, 128, smem>>>(dA, dB, dD);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxerr_hi = 0;
    for (int m = 0; m < 40; ++m) {
        const int lane = (m / 16) * 32 + m % 16;
        for (int n = 0; n < 40; ++n) {
            double r = 0;
            for (int k = 0; k < 128; ++k) r += (double)A[k * 40 + m] * (double)B[k * 40 + n];
            const double e = fabs(r - D[lane * 40 + n]);
            if (m < 32 && n < 32) maxerr = fmax(maxerr, e); else maxerr_hi = fmax(maxerr_hi, e);
        }
    }
    printf("shared second atom: max abs err channels < 32: %.3e, channels 32..39 (rows or columns): %.3e  %s\n", maxerr, maxerr_hi,
           (maxerr < 5e-5 && maxerr_hi < 5e-5) ? "OK" : "FAIL");
    return 0;
}
