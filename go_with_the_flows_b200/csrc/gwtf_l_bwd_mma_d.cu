// mma.sync register-fragment backward, phase 0 (gwtf_bwd_mma.cuh: k_bwd_layer_d_mma).
#include "gwtf_host.h"
#include "gwtf_bwd_mma.cuh"

namespace gwtf {

template <int FP>
static int launch_d(const BwdArgs& a, cudaStream_t st) {
    const int F = a.d.n_features, K = a.d.n_components;
    const size_t smem = bwd_d_mma_smem<FP>(F);
    auto kern = k_bwd_layer_d_mma<FP>;
    GWTF_CUDA(allow_smem(kern, smem));
    const long long tiles = (long long)a.B * ((a.N + 127) / 128);
    int gx = BwdDTmem<FP>::ctas_per_sm * num_sms() / K;      // tensor-memory columns bound the residency; contiguous tile ranges
    if (gx > tiles) gx = (int)tiles;
    GWTF_CUDA(launch_pdl(pdl_on(a.d), kern, dim3(gx < 1 ? 1 : gx, K), dim3(kThreads), smem, st, a));
    return 0;
}

int launch_bwd_layer_d_mma(const BwdArgs& a, cudaStream_t st) {
    GWTF_DISPATCH_FP8(a.d.n_features, return launch_d<FP>(a, st));
    return 0;
}

int launch_bwd_layer_mma(const BwdArgs& a, int phase, cudaStream_t st) {
    return phase == 0 ? launch_bwd_layer_d_mma(a, st) : launch_bwd_layer_e_mma(a, st);
}

}  // namespace gwtf
