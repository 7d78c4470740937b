// mma.sync register-fragment backward, phase 1 (gwtf_bwd_mma.cuh: k_bwd_layer_e_mma).
#include "gwtf_host.h"
#include "gwtf_bwd_mma.cuh"

namespace gwtf {

template <int FP>
static int launch_e(const BwdArgs& a, cudaStream_t st) {
    const int F = a.d.n_features, K = a.d.n_components;
    const size_t smem = bwd_e_mma_smem<FP>(F);
    auto kern = k_bwd_layer_e_mma<FP>;
    GWTF_CUDA(allow_smem(kern, smem));
    const long long tiles = (long long)a.B * ((a.N + 127) / 128);
    int per_sm = 512 / BwdETmem<FP>::alloc;                  // tensor-memory columns bound the residency
    if (per_sm > 2) per_sm = 2;
    int gx = per_sm * num_sms() / K;                         // contiguous tile ranges
    if (gx > tiles) gx = (int)tiles;
    GWTF_CUDA(launch_pdl(pdl_on(a.d), kern, dim3(gx < 1 ? 1 : gx, K), dim3(kThreads), smem, st, a));
    return 0;
}

int launch_bwd_layer_e_mma(const BwdArgs& a, cudaStream_t st) {
    GWTF_DISPATCH_FP8(a.d.n_features, return launch_e<FP>(a, st));
    return 0;
}

}  // namespace gwtf
