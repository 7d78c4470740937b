// tcgen05 / TMEM backward of one coupling layer (gwtf_tc_bwd.cuh).
#include "gwtf_host.h"
#include "gwtf_tc_bwd.cuh"

namespace gwtf {

template <int FPK, int FPN, int PHASE>
static int launch_bwd_tc(const BwdArgs& a0, cudaStream_t st) {
    BwdArgs a = a0;
    const int F = a.d.n_features, K = a.d.n_components;
    const size_t smem = bwd_tc_smem<FPK, FPN>(F);
    auto kern = k_bwd_layer_tc<FPK, FPN, PHASE>;
    GWTF_CUDA(allow_smem(kern, smem));
    const int tiles = a.B * ((a.N + 127) / 128);
    int gx = num_sms() / K;                                  // one CTA per SM owns all 512 TMEM columns
    if (gx > (tiles + kBSlots - 1) / kBSlots) gx = (tiles + kBSlots - 1) / kBSlots;
    if (gx < 1) gx = 1;
    GWTF_CUDA(launch_pdl(pdl_on(a.d), kern, dim3(gx, K), dim3(kBwdThreads), smem, st, a));
    return 0;
}

int launch_bwd_layer_tc(const BwdArgs& a, int phase, cudaStream_t st) {
    if (a.d.n_components > num_sms()) return fail(-4, "more components than SMs");
    if (phase == 0) return launch_bwd_layer_d_mma(a, st);          // phase 0: register-fragment kernel (recomputing)
    GWTF_DISPATCH_TC(a.d.n_features, return (launch_bwd_tc<FPK, FPN, 1>(a, st)));
    return 0;
}

}  // namespace gwtf
