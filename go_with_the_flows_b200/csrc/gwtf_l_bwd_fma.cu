// FP32-FMA backward phases + the engine-independent backward kernels (seeds, closed-form finish).
#include "gwtf_host.h"
#include "gwtf_bwd.cuh"

namespace gwtf {

template <int FP, int PHASE>
static int launch_bwd_layer_t(const BwdArgs& a0, cudaStream_t st) {
    constexpr int P = PHASE == 0 ? PointsPerThread<FP>::fwd : PointsPerThread<FP>::bwd;
    BwdArgs a = a0;
    a.tiles_per_shape = (a.N + kThreads * P - 1) / (kThreads * P);
    const int F = a.d.n_features, K = a.d.n_components;
    const int tiles = a.B * a.tiles_per_shape;
    if constexpr (PHASE == 0) {
        const size_t smem = round_up((int)sizeof(BwdSmem<FP>), 16) + (size_t)round_up(raw_floats(F), 4) * 4 +
                            (size_t)FP * P * kThreads * 4;
        auto kern = k_bwd_layer_d<FP, P>;
        GWTF_CUDA(allow_smem(kern, smem));
        int gx = (num_sms() * blocks_per_sm(kern, smem) + K - 1) / K;
        if (gx > tiles) gx = tiles;
        kern<<<dim3(gx < 1 ? 1 : gx, K), kThreads, smem, st>>>(a);
    } else {
        const size_t smem = round_up((int)sizeof(BwdESmem<FP>), 16) + (size_t)round_up(raw_floats(F), 4) * 4 +
                            2 * (size_t)P * kThreads * FP * 4 + 256;     // + slack: MMA fragments read past row ends
        auto kern = k_bwd_layer_e<FP, P>;
        GWTF_CUDA(allow_smem(kern, smem));
        int gx = (num_sms() * blocks_per_sm(kern, smem) + K - 1) / K;
        if (gx > tiles) gx = tiles;
        kern<<<dim3(gx < 1 ? 1 : gx, K), kThreads, smem, st>>>(a);
    }
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

int launch_bwd_layer_fma(const BwdArgs& a, int phase, cudaStream_t st) {
    if (phase == 0) { GWTF_DISPATCH_FP(a.d.n_features, return (launch_bwd_layer_t<FP, 0>(a, st))); }
    else { GWTF_DISPATCH_FP(a.d.n_features, return (launch_bwd_layer_t<FP, 1>(a, st))); }
    return 0;
}

int launch_bwd_seed(int K, int B, int N, const float* ubuf, const float* ld, const float* base, const float* logw,
                    const float* nll, const float* dnll, float* gbuf, float* gs, float* dbase, float* dlogw,
                    cudaStream_t st) {
    int gx = (N + kThreads - 1) / kThreads;
    if (gx > 64) gx = 64;
    k_bwd_seed<<<dim3(gx, B), kThreads, 0, st>>>(K, B, N, ubuf, ld, base, logw, nll, dnll, gbuf, gs, dbase, dlogw);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

template <int FP>
static int launch_finish_points(const gwtf_stack_desc& d, int train, const float* params, const double* mom_last,
                                const double* bsum_last, const float* gbuf, const float* points, float* dpoints, int B,
                                int N, double n_total, cudaStream_t st) {
    const size_t total = (size_t)B * N;
    int grid = (int)((total + kThreads - 1) / kThreads);
    if (grid > 8 * num_sms()) grid = 8 * num_sms();
    k_bwd_finish_points<FP><<<grid, kThreads, 0, st>>>(d, train, params, mom_last, bsum_last, gbuf, points, dpoints, B, N,
                                                       n_total);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

int launch_bwd_finish(const FinishArgs& fa, const gwtf_stack_desc& d, int train, const float* params, const double* mom,
                      const double* bsum, const float* gbuf, const float* points, float* dpoints, int B, int N,
                      double n_total, cudaStream_t st) {
    const int L = d.n_layers, K = d.n_components, F = d.n_features;
    const int total = L * K * 2 * F;
    k_bwd_finish_w0<<<(total + 127) / 128, 128, 0, st>>>(fa);
    GWTF_CUDA(cudaGetLastError());
    if (dpoints) {
        const double* mom_last = train ? mom + (size_t)(L - 1) * K * GWTF_MOM_STRIDE : nullptr;
        const double* bsum_last = train ? bsum + (size_t)(L - 1) * K * 8 * F : nullptr;
        GWTF_DISPATCH_FP(F, return launch_finish_points<FP>(d, train, params, mom_last, bsum_last, gbuf, points, dpoints, B,
                                                            N, n_total, st));
    }
    return 0;
}

}  // namespace gwtf
