// FP32-FMA forward kernels + the small engine-independent kernels (moments, bstat, mixture head).
#include "gwtf_host.h"
#include "gwtf_fwd.cuh"

namespace gwtf {

template <int FP>
static int launch_eval_t(const EvalArgs& a0, cudaStream_t st) {
    constexpr int P = PointsPerThread<FP>::fwd;
    EvalArgs a = a0;
    a.tiles_per_shape = (a.N + kThreads * P - 1) / (kThreads * P);
    const int F = a.d.n_features;
    const size_t smem = round_up((int)sizeof(EvalSmem<FP>), 16) + 2 * (size_t)round_up(raw_floats(F), 4) * 4;
    GWTF_CUDA(allow_smem(k_nll_eval<FP, P>, smem));
    k_nll_eval<FP, P><<<a.B * a.tiles_per_shape, kThreads, smem, st>>>(a);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

int launch_eval_fma(const EvalArgs& a, cudaStream_t st) {
    GWTF_DISPATCH_FP(a.d.n_features, return launch_eval_t<FP>(a, st));
    return 0;
}

template <int FP, int PHASE>
static int launch_fwd_layer_t(const LayerArgs& a0, cudaStream_t st) {
    constexpr int P = PointsPerThread<FP>::fwd;
    LayerArgs a = a0;
    a.tiles_per_shape = (a.N + kThreads * P - 1) / (kThreads * P);
    const int F = a.d.n_features;
    const size_t smem = round_up((int)sizeof(PhaseSmem<FP>), 16) + (size_t)round_up(raw_floats(F), 4) * 4;
    auto kern = k_fwd_layer<FP, P, PHASE>;
    GWTF_CUDA(allow_smem(kern, smem));
    const int tiles = a.B * a.tiles_per_shape;
    const int K = a.d.n_components;
    int gx = (num_sms() * blocks_per_sm(kern, smem) + K - 1) / K;
    if (gx > tiles) gx = tiles;
    if (gx < 1) gx = 1;
    kern<<<dim3(gx, K), kThreads, smem, st>>>(a);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

int launch_fwd_layer_fma(const LayerArgs& a, int phase, cudaStream_t st) {
    if (phase == 0) { GWTF_DISPATCH_FP(a.d.n_features, return (launch_fwd_layer_t<FP, 0>(a, st))); }
    else { GWTF_DISPATCH_FP(a.d.n_features, return (launch_fwd_layer_t<FP, 1>(a, st))); }
    return 0;
}

int launch_moments(const float* points, int B, int N, int K, double* mom, cudaStream_t st) {
    size_t total = (size_t)B * N;
    int grid = (int)((total + kThreads * 8 - 1) / (kThreads * 8));
    if (grid > 4 * num_sms()) grid = 4 * num_sms();
    k_moments<<<grid, kThreads, 0, st>>>(points, B, N, K, mom);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

int launch_bstat(const gwtf_stack_desc& d, const float* params, const double* mom, const double* sum1, double n_total,
                 float* bstat, cudaStream_t st) {
    const int total = d.n_layers * d.n_components * 2 * d.n_features;
    k_bstat<<<(total + 255) / 256, 256, 0, st>>>(d, params, mom, sum1, n_total, bstat);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

int launch_nll_from_state(const gwtf_stack_desc& d, int B, int N, const float* ubuf, const float* ld, const float* base,
                          const float* logw, float* nll, float* logp, cudaStream_t st) {
    const size_t total = (size_t)B * N;
    k_nll_from_state<<<(unsigned)((total + kThreads - 1) / kThreads), kThreads, 0, st>>>(d.n_components, B, N, ubuf, ld,
                                                                                        base, logw, nll, logp, d.nonfinite);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gwtf
