// Sampling kernels: the fused FP32-FMA sampler and the draw / regroup / gather kernels around the layer kernels.
#include "gwtf_host.h"
#include "gwtf_sample.cuh"

namespace gwtf {

template <int FP>
static int launch_sample_t(const SampleArgs& a0, cudaStream_t st) {
    constexpr int PMAX = PointsPerThread<FP>::fwd;
    SampleArgs a = a0;
    // tile size: large tiles amortise the K*L record stream, small ones fill the SMs
    int tile = kSampleMaxTile;
    while (tile > kThreads && (long long)a.B * ((a.N + tile - 1) / tile) < 2LL * num_sms()) tile >>= 1;
    a.tile_points = tile;
    a.tiles_per_shape = (a.N + tile - 1) / tile;
    const int F = a.d.n_features;
    const size_t smem = round_up((int)sizeof(SampleSmem<FP>), 16) + 2 * (size_t)round_up(raw_floats(F), 4) * 4;
    auto kern = k_sample<FP, PMAX>;
    GWTF_CUDA(allow_smem(kern, smem));
    kern<<<a.B * a.tiles_per_shape, kThreads, smem, st>>>(a);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

int launch_sample_fma(const SampleArgs& a, cudaStream_t st) {
    GWTF_DISPATCH_FP(a.d.n_features, return launch_sample_t<FP>(a, st));
    return 0;
}

int launch_mixture_cdf(const float* logits, int B, int K, float* cdf, cudaStream_t st) {
    k_mixture_cdf<<<(B + 127) / 128, 128, 0, st>>>(logits, B, K, cdf);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

int launch_sample_count(const SamplePlanArgs& a, cudaStream_t st) {
    k_sample_count<<<dim3((a.N + kSampleSpan - 1) / kSampleSpan, a.B), kThreads, 0, st>>>(a);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

int launch_sample_plan(int K, int B, const int32_t* counts, int32_t* seg, int32_t* seg_tiles, int32_t* cursor,
                       cudaStream_t st) {
    k_sample_plan<<<1, kThreads, 0, st>>>(K, B, counts, seg, seg_tiles, cursor);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

int launch_sample_scatter(const SampleScatterArgs& a, cudaStream_t st) {
    k_sample_scatter<<<dim3((a.N + kSampleSpan - 1) / kSampleSpan, a.B), kThreads, 0, st>>>(a);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

int launch_sample_gather(int B, int N, int Npad, const float* xin, const int32_t* slot, float* samples, cudaStream_t st) {
    const size_t total = (size_t)B * N;
    int grid = (int)((total + kThreads - 1) / kThreads);
    if (grid > 8 * num_sms()) grid = 8 * num_sms();
    k_sample_gather<<<grid, kThreads, 0, st>>>(B, N, Npad, xin, slot, samples);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gwtf
