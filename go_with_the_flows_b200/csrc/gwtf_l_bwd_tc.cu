// tcgen05 / TMEM backward of one coupling layer (gwtf_tc_bwd.cuh).
#include "gwtf_host.h"
#include "gwtf_tc_bwd.cuh"

namespace gwtf {

template <int FPK, int FPN, int PHASE>
static int launch_bwd_tc(const BwdArgs& a0, cudaStream_t st) {
    BwdArgs a = a0;
    const int F = a.d.n_features, K = a.d.n_components;
    constexpr int SLOTS = PHASE == 0 ? kDSlots : kBSlots;
    constexpr int THREADS = PHASE == 0 ? kBwdDThreads : kBwdThreads;
    const size_t smem = PHASE == 0 ? bwd_tc_d_smem<FPK, FPN>(F) : bwd_tc_smem<FPK, FPN>(F);
    auto kern = k_bwd_layer_tc<FPK, FPN, PHASE>;
    GWTF_CUDA(allow_smem(kern, smem));
    const int tiles = a.B * ((a.N + 127) / 128);
    int gx = num_sms() / K;                                  // one CTA per SM owns all 512 TMEM columns
    if (gx > (tiles + SLOTS - 1) / SLOTS) gx = (tiles + SLOTS - 1) / SLOTS;
    if (gx < 1) gx = 1;
    GWTF_CUDA(launch_pdl(pdl_on(a.d), kern, dim3(gx, K), dim3(THREADS), smem, st, a));
    return 0;
}

int launch_bwd_layer_tc(const BwdArgs& a, int phase, cudaStream_t st) {
    if (a.d.n_components > num_sms()) return fail(-4, "more components than SMs");
    if (phase == 0) {
        if (getenv("GWTF_BWD_D_MMA")) return launch_bwd_layer_d_mma(a, st);      // (bring-up switch: register-fragment phase 0)
        GWTF_DISPATCH_TC(a.d.n_features, return (launch_bwd_tc<FPK, FPN, 0>(a, st)));
    }
    GWTF_DISPATCH_TC(a.d.n_features, return (launch_bwd_tc<FPK, FPN, 1>(a, st)));
    return 0;
}

}  // namespace gwtf

// Host-side replay of the tile schedule of the persistent layer kernels (the same RoundIter / SlotTurns code the
// kernels run): which CTA, tile slot and per-CTA sequence number every 128-point tile of a launch gets.
extern "C" int gwtf_debug_tile_schedule(int32_t total_tiles, int32_t tiles_per_shape, int32_t grid_x, int32_t slots,
                                        int32_t* tile_cta, int32_t* tile_slot, int32_t* tile_seq) {
    using namespace gwtf;
    if (total_tiles < 0 || tiles_per_shape < 1 || grid_x < 1 || slots < 1 || !tile_cta || !tile_slot || !tile_seq)
        return fail(-10, "bad tile schedule query");
    const int B = (total_tiles + tiles_per_shape - 1) / tiles_per_shape;
    const int per_cta = (total_tiles + grid_x - 1) / grid_x;
    int assigned = 0;
    for (int cta = 0; cta < grid_x; ++cta) {
        const int t_begin = min(cta * per_cta, total_tiles), t_end = min(t_begin + per_cta, total_tiles);
        for (int s = 0; s < slots; ++s) {
            RoundIter it(t_begin, t_end, tiles_per_shape, nullptr, B, slots);
            SlotTurns turns{0};
            int base, count, b, seq = 0;
            while (it.next(base, count, b)) {
                int off = s, my_seq = seq + s;
                bool mine = s < count;
                if (slots == kBSlots) mine = turns.take(s, count, off, my_seq);        // backward phase 1: strict alternation
                seq += count;
                if (!mine) continue;
                const int t = base + off;
                if (t < t_begin || t >= t_end || t / tiles_per_shape != b) return fail(-27, "tile outside its range or shape");
                tile_cta[t] = cta; tile_slot[t] = s; tile_seq[t] = my_seq;
                ++assigned;
            }
        }
    }
    return assigned;
}

#ifdef GWTF_STAGE_CLOCKS
extern "C" int gwtf_debug_stage_clocks(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    if (out) cudaMemcpyFromSymbol(out, gwtf::g_stage_clk, sizeof(gwtf::g_stage_clk));
    if (reset) { unsigned long long z[32] = {}; cudaMemcpyToSymbol(gwtf::g_stage_clk, z, sizeof(z)); }
    return 0;
}
#endif
