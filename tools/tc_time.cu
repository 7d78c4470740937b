// Cycle breakdown of the tcgen05 forward layer kernel on a C2-sized synthetic layer.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DGWTF_TIMING -o tools/tc_time tools/tc_time.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../go_with_the_flows_b200/csrc/gwtf_tc_fwd.cuh"
using namespace gwtf;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int PHASE, int TRAIN>
int run(int per_sm) {
    const int K = 4, L = 33, F = 37, B = 64, N = 2048;
    gwtf_stack_desc d{};
    d.n_components = K; d.n_layers = L; d.n_features = F; d.rec_stride = rec_stride_of(F);
    for (int l = 0; l < L; ++l) d.warp_mask[l] = ((l / 3) % 2 == 0) ? (1 << (l % 3)) : (7 ^ (1 << (2 - l % 3)));
    std::vector<float> hp((size_t)K * L * d.rec_stride), hf((size_t)B * K * L * 4 * F), hx((size_t)K * B * 3 * N), hbn((size_t)K * L * 8 * F, 1.0f);
    for (auto& v : hp) v = 0.3f * ((float)rand() / RAND_MAX - 0.5f);
    for (auto& v : hf) v = 0.5f + (float)rand() / RAND_MAX;
    for (auto& v : hx) v = 0.4f * ((float)rand() / RAND_MAX - 0.5f);
    float *p, *f, *x, *bn, *xo, *ss; double *mom, *sum1;
    CK(cudaMalloc(&p, hp.size() * 4)); CK(cudaMalloc(&f, hf.size() * 4)); CK(cudaMalloc(&x, hx.size() * 4)); CK(cudaMalloc(&bn, hbn.size() * 4));
    CK(cudaMalloc(&xo, hx.size() * 4)); CK(cudaMalloc(&ss, hx.size() * 4));
    CK(cudaMalloc(&mom, K * 16 * 8)); CK(cudaMalloc(&sum1, K * 4 * F * 8));
    double* mom_in; CK(cudaMalloc(&mom_in, K * 16 * 8));
    { std::vector<double> hm(K * 16, 0.0); for (int j = 0; j < K; ++j) { double n = (double)B * N; hm[j*16+3] = 0.05 * n; hm[j*16+6] = 0.04 * n; hm[j*16+8] = 0.03 * n; }
      CK(cudaMemcpy(mom_in, hm.data(), hm.size() * 8, cudaMemcpyHostToDevice));
      std::vector<double> hs(K * 4 * F); for (int i = 0; i < K * 4 * F; ++i) hs[i] = ((i / F) % 2) ? 0.5 * B * N : 0.01 * B * N; CK(cudaMemcpy(sum1, hs.data(), hs.size() * 8, cudaMemcpyHostToDevice)); }
    CK(cudaMemcpy(p, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(f, hf.data(), hf.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(x, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(bn, hbn.data(), hbn.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(ss, 0, hx.size() * 4)); CK(cudaMemset(mom, 0, K * 16 * 8));
    LayerArgs a{};
    a.d = d; a.layer = 5; a.train = TRAIN; a.direct = 0; a.params = p; a.bnbuf = bn; a.film = f; a.xin = x; a.xin_shared = 0;
    a.xout = xo; a.ld = nullptr; a.ssum = ss; a.trio = nullptr; a.mom_in = mom_in; a.mom_out = PHASE == 1 ? mom : nullptr; a.sum1 = sum1;
    a.B = B; a.N = N; a.tiles_per_shape = N / 128; a.n_total = (double)B * N;
    const size_t smem = round_up((int)sizeof(TcFwdSmem<40, 48>), 16) + (size_t)round_up(raw_floats(F), 4) * 4;
    auto kern = k_fwd_layer_tc<40, 48, PHASE>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int gx = (148 * per_sm + K - 1) / K;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<dim3(gx, K), 128, smem>>>(a);
    CK(cudaDeviceSynchronize());
    long long zero[16] = {0};
    CK(cudaMemcpyToSymbol(g_tc_cycles, zero, sizeof(zero)));
    cudaEventRecord(e0);
    kern<<<dim3(gx, K), 128, smem>>>(a);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c[16]; CK(cudaMemcpyFromSymbol(c, g_tc_cycles, sizeof(c)));
    const int tiles_cta0 = (B * (N / 128) + gx - 1) / gx;
    printf("train %d phase %d  %d CTA/SM grid %dx%d: %.1f us; CTA0 (%d tiles) cycles: tma %lld vectors %lld operands %lld | restage_b1 %lld x+handoff %lld to_h1 %lld sums %lld relu2 %lld mma2 %lld read_o %lld tail %lld\n",
           TRAIN, PHASE, per_sm, gx, K, ms * 1e3, tiles_cta0, c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8], c[9], c[10]);
    return 0;
}
int main() {
    for (int per_sm : {3}) { if (run<0, 0>(per_sm)) return 1; if (run<1, 0>(per_sm)) return 1; if (run<0, 1>(per_sm)) return 1; if (run<1, 1>(per_sm)) return 1; }
    return 0;
}
