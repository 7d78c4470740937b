// mma.sync register-fragment forward of one coupling layer (gwtf_fwd_mma.cuh): feature widths up to 64.
#include "gwtf_host.h"
#include "gwtf_fwd.cuh"
#include "gwtf_fwd_mma.cuh"

namespace gwtf {

template <int FP, int PHASE>
static int launch_fwd_layer_mma_t(const LayerArgs& a, cudaStream_t st) {
    // statistics pass: 2 m-tiles per warp (shared B fragments, 2 MMA chains), full register file, 1 CTA/SM;
    // apply pass: 1 m-tile per warp, 128 registers, 2 CTAs/SM
    constexpr int MI = PHASE == 0 ? 2 : 1;
    const int F = a.d.n_features, K = a.d.n_components;
    const size_t smem = fwd_mma_smem<FP>(F);
    auto kern = k_fwd_layer_mma<FP, PHASE, MI>;
    GWTF_CUDA(allow_smem(kern, smem));
    const long long tiles = (long long)a.B * ((a.N + 128 * MI - 1) / (128 * MI));
    int gx = (MI == 1 ? 2 : 1) * num_sms() / K;              // contiguous tile ranges
    if (gx > tiles) gx = (int)tiles;
    kern<<<dim3(gx < 1 ? 1 : gx, K), kThreads, smem, st>>>(a, PHASE == 1 ? a.y1out : nullptr);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

int launch_fwd_layer_mma(const LayerArgs& a, int phase, cudaStream_t st) {
    if (a.seg) return fail(-4, "segmented rows need the tcgen05 forward");
    if (phase == 0) { GWTF_DISPATCH_FP8(a.d.n_features, return (launch_fwd_layer_mma_t<FP, 0>(a, st))); }
    else { GWTF_DISPATCH_FP8(a.d.n_features, return (launch_fwd_layer_mma_t<FP, 1>(a, st))); }
    return 0;
}

}  // namespace gwtf
