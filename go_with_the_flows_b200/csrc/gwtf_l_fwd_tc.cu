// tcgen05 / TMEM forward of one coupling layer (persistent warp-specialised kernel, gwtf_tc_persist.cuh).
#include "gwtf_host.h"
#include "gwtf_tc_persist.cuh"

namespace gwtf {

template <int FPK, int FPN, int PHASE, int PASSES>
static int launch_fwd_layer_tcp(const LayerArgs& a0, cudaStream_t st) {
    LayerArgs a = a0;
    a.tiles_per_shape = (a.N + 127) / 128;
    const int F = a.d.n_features, K = a.d.n_components;
    const size_t smem = round_up((int)sizeof(TcPersistSmem<FPK, FPN>), 16) + (size_t)round_up(raw_floats(F), 4) * 4;
    auto kern = k_fwd_layer_tcp<FPK, FPN, PHASE, PASSES>;
    GWTF_CUDA(allow_smem(kern, smem));
    const int tiles = a.B * a.tiles_per_shape;               // (segmented mode: an upper bound)
    int gx = num_sms() / K;                                  // one CTA per SM owns all 512 TMEM columns
    if (gx > (tiles + kSlots - 1) / kSlots) gx = (tiles + kSlots - 1) / kSlots;
    if (gx < 1) gx = 1;
    GWTF_CUDA(launch_pdl(pdl_on(a.d), kern, dim3(gx, K), dim3(kPersistThreads), smem, st, a));
    return 0;
}

// phase 0 = train-mode statistics, 1 = apply.  Single-pass TF32 only where the caller asked for it AND nothing is
// differentiated through the result: eval-mode apply passes (a.train == 0) with eval_precision set.
int launch_fwd_layer_tc(const LayerArgs& a, int phase, cudaStream_t st) {
    const int F = a.d.n_features;
    if (a.d.n_components > num_sms()) return fail(-4, "more components than SMs");
    const bool fast = phase == 1 && !a.train && a.d.eval_precision == GWTF_PRECISION_TF32 && !a.y1out;
    if (phase == 0) { GWTF_DISPATCH_TC(F, return (launch_fwd_layer_tcp<FPK, FPN, 0, 3>(a, st))); }
    else if (fast) { GWTF_DISPATCH_TC(F, return (launch_fwd_layer_tcp<FPK, FPN, 1, 1>(a, st))); }
    else { GWTF_DISPATCH_TC(F, return (launch_fwd_layer_tcp<FPK, FPN, 1, 3>(a, st))); }
    return 0;
}

}  // namespace gwtf
