// Backward of one coupling layer on the tcgen05 tensor cores (TMEM accumulators), same persistent warp-specialised
// structure as the forward (gwtf_tc_persist.cuh): compute warpgroups (128 threads = 128 TMEM lanes = one 128-point
// tile in flight each) + one MMA-issuer warp per tile slot, mbarrier hand-offs in both directions.
// Same BwdArgs and outputs as the mma.sync kernels (gwtf_bwd_mma.cuh); math: SURVEY.md App. F.
//
// Phase 1 (k_bwd_layer_tc<.., 1>), per 128-point tile and net -- every contraction is a UMMA chain, 3xTF32:
//   batch A   y0 = X B0^T                  (TS, K = 8: the (x,1 | dO) operand row of every point, written to the TMEM
//                                           columns a0 takes over afterwards: no shared-memory write, no proxy fence)
//             P  = X PW^T                  dO (alpha W2): the sd2 backward with s/sigma1 folded in
//   C0        a0 = relu(y0) -> TMEM operand (hi | lo), [y0 > 0] kept as a bit mask
//   batch B   y1 = a0 B1^T                 (TS) sd1 + BN1 + FiLM folded into B1 (bias column): y1 leaves finished
//   C1        r = [y1 > 0] P + beta y1 + gamma        = dh1 (BN1 backward is affine in y1; beta = gamma = 0 in eval)
//             r -> TMEM operand AND, with a0, -> the point-contraction operands in shared memory (MN-major)
//   batch C   da0 = r W1                   (TS)
//             dW1 += r^T a0                (SS, K = the tile's 128 points, M = 64, accumulator resident in TMEM for the
//                                           whole kernel; committed separately so the tile goes on while it runs)
//   C2        dy0 = [y0 > 0] da0 -> TMEM operand; per-channel sums of dy0 (1, x_keep) as column sums through the warp's
//             shared-memory tile (gwtf_tc_fwd.cuh: col_write_row)
//   batch D   du = dy0 Q0^T                (TS, N = 16): the input gradient
// The two nets are processed one after the other (outer loop) so that one net's operands fit shared memory next to
// the 80 KB of point-contraction operands, which the tile slots take turns on (one mbarrier; the slots strictly
// alternate, see SlotTurns).
#pragma once
#include "gwtf_tc_persist.cuh"
#include "gwtf_bwd.cuh"

namespace gwtf {

constexpr int kBSlots = 2;
constexpr int kBwdThreads = kBSlots * 128 + kBSlots * 32;      // compute warpgroups + one issuer warp per slot
// Point-contraction operands (MN-major, 128B swizzle / 32B base): an mn-atom is 32 channels x 128 points = 16 KB.  Each of
// the four arrays (r_hi, r_lo, a0_hi, a0_lo) owns ONE atom (channels 0..31); channels 32..39 of array `a` live in 32-byte
// slot `a` of every 128-byte row of a fifth, shared atom, which each array's descriptor reaches through its own
// leading-dimension offset (tools/tc_probe5.cu): 80 KB instead of 128 KB.  Rows 40..63 of the M = 64 operand read
// whatever the neighbouring slots hold -- they only feed accumulator rows nobody reads.
constexpr int kMnAtom = (128 / 4) * 128;                       // floats per atom
constexpr int kMnFloats = kMnAtom;                             // stride between the four arrays
constexpr int kMnTotal = 5 * kMnAtom + 64;                     // + the shared atom + slack for reads past its end

__device__ __forceinline__ void bwd_compute_barrier() { asm volatile("bar.sync 2, %0;" ::"n"(kBSlots * 128) : "memory"); }
__device__ __forceinline__ void wg_barrier(int slot) { asm volatile("bar.sync %0, 128;" ::"r"(3 + slot) : "memory"); }

template <int FPK, int FPN>
struct TcBwdECols {               // tensor-memory columns of one tile slot
    static constexpr uint32_t D = 0, P = FPN, Ahi = 2 * FPN, Alo = 2 * FPN + FPK, SLOT = 2 * FPN + 2 * FPK;
    static constexpr uint32_t G = kBSlots * SLOT;              // dW1 accumulators: FPK columns per slot (M = 64 rows)
    static_assert(kBSlots * (SLOT + FPK) <= 512, "tensor memory budget");
};

template <int FPK, int FPN>
struct TcBwdESmem {
    LayerT<FPN> W;
    LayerWB<FPN> WB;
    TcOperand<FPN, 8> B0;          // [e][x0,x1,x2,1,0,0,0,0]        BN0-folded sd0 (+ bias); row F makes the constant one
    TcOperand<FPN, 8> PW;          // [f][0,0,0,0,c2x,c2y,c2z,0]     c2 = (s/sigma1) * sd2 columns        (per shape)
    TcOperand<FPN, FPK> B1;        // [f][e] folded sd1 | bias column                                       (per shape)
    TcOperand<FPN, FPK> W1T;       // [e][f] = W1[f][e]
    TcOperand<16, FPK> Q0;         // [d][e] = q0[e].d
    float2 bg[FPN];                // (beta, gamma) of the current shape
    float red[4 * 32];
    alignas(16) float col[kBSlots * 4][kColTile];
    uint64_t bar_tma, req[kBSlots], done[kBSlots], buf_free;
    uint32_t tmem_base;
};

// D[M=64 x N] (+)= A^T B over the 128 points of a tile, both operands MN-major (sw32) in shared memory, 3xTF32
__device__ __forceinline__ uint64_t mn_operand_desc(const float* mn, int arr) {
    const float* base = mn + arr * kMnFloats;
    const uint32_t lbo = smem_u32(mn + 4 * kMnFloats + 8 * arr) - smem_u32(base);      // -> slot `arr` of the shared atom
    return make_smem_desc(base, lbo, 512u, 1u);
}
template <int N>
__device__ __forceinline__ void issue_point_contraction(uint32_t d_tmem, const float* mn) {
    const uint32_t idesc = make_idesc_tf32(64, N, 1, 1);
    uint64_t ah = mn_operand_desc(mn, 0), al = mn_operand_desc(mn, 1);
    uint64_t bh = mn_operand_desc(mn, 2), bl = mn_operand_desc(mn, 3);
    // opaque to the optimiser: otherwise all 64 per-step descriptors are hoisted out of the tile loop as loop invariants
    // and spilled (local memory misses the small L1 here), putting a memory round trip in front of every MMA issue
    asm volatile("" : "+l"(ah), "+l"(al), "+l"(bh), "+l"(bl));
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint64_t a = pass == 1 ? al : ah;
        const uint64_t b = pass == 2 ? bl : bh;
#pragma unroll
        for (int s = 0; s < 16; ++s)
            mma_tf32_ss(d_tmem, a + (uint64_t)((s * 1024) >> 4), b + (uint64_t)((s * 1024) >> 4), idesc, true);
    }
}

// this thread's 8 channels (one 32-byte chunk) of point `p` into an MN-major operand
__device__ __forceinline__ void store_mn_chunk(float* mn, int arr, int chunk, int p, const float (&v)[8]) {
    float* dst = chunk < 4 ? mn + arr * kMnFloats + mnmajor_sw32_offset(8 * chunk, p, 128)
                           : mn + 4 * kMnFloats + (p >> 2) * 128 + (p & 3) * 32 + ((arr ^ (p & 3)) << 3);
    // The 8 lanes that share p & 3 land in the same 32-byte bank window (rows 512 B apart): half of them store the upper
    // 16 bytes first, so each instruction covers all 32 banks (4 wavefronts instead of 8).
    const bool odd = (p >> 2) & 1;
    const float4 a = make_float4(v[0], v[1], v[2], v[3]), b = make_float4(v[4], v[5], v[6], v[7]);
    *reinterpret_cast<float4*>(dst + (odd ? 4 : 0)) = odd ? b : a;
    *reinterpret_cast<float4*>(dst + (odd ? 0 : 4)) = odd ? a : b;
}

// Stage clocks (profiling builds only: GWTF_NVCC_EXTRA=-DGWTF_STAGE_CLOCKS, tools/stage_clocks.py): one compute thread
// of CTA (0, 0) accumulates the cycles it spends between consecutive marks.
#ifdef GWTF_STAGE_CLOCKS
__device__ unsigned long long g_stage_clk[32];
__shared__ unsigned int s_stage_clk[32];
#define GWTF_CLK_INIT(cond) const bool clk_on = (cond); unsigned int clk_last = (unsigned int)clock();
#define GWTF_CLK(i) if (clk_on) { const unsigned int now_ = (unsigned int)clock(); s_stage_clk[i] += now_ - clk_last; clk_last = now_; }
#define GWTF_CLK_FLUSH() { __syncthreads(); if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 32) { g_stage_clk[threadIdx.x] += s_stage_clk[threadIdx.x]; } }
#define GWTF_CLK_ZERO() { if (threadIdx.x < 32) s_stage_clk[threadIdx.x] = 0u; }
#else
#define GWTF_CLK_INIT(cond)
#define GWTF_CLK(i)
#define GWTF_CLK_FLUSH()
#define GWTF_CLK_ZERO()
#endif

// Which slot takes which tile of a round (phase 1, two slots).  The point contractions of a CTA share one operand buffer
// and are numbered in processing order; contraction k waits for the completion of k - 1 on an mbarrier by phase parity,
// which is only unambiguous if the two slots strictly alternate.  Rounds can hold a single tile (a tile range starting at
// an odd offset in its shape, a shape with an odd tile count), so the round's first tile goes to the slot whose turn it
// is -- slot (seq & 1) -- not always to slot 0.
struct SlotTurns {
    int seq;                    // contractions of this CTA before the current round
    // slot s in a round of `count` tiles: does it have a tile, which (offset in the round), and its sequence number
    __host__ __device__ __forceinline__ bool take(int s, int count, int& off, int& my_seq) {
        off = (s - seq) & 1;
        my_seq = seq + off;
        seq += count;
        return off < count;
    }
};

template <int FPK, int FPN>
__device__ __forceinline__ void bwd_tc_phase1(const BwdArgs& a, unsigned char* smem_raw) {
    using SM = TcBwdESmem<FPK, FPN>;
    using C = TcBwdECols<FPK, FPN>;
    SM& S = *reinterpret_cast<SM*>(smem_raw);
    float* raw = reinterpret_cast<float*>(smem_raw + round_up((int)sizeof(SM), 16));
    const int F = a.d.n_features, K = a.d.n_components, L = a.d.n_layers;
    // the point-contraction operands, 1024-byte aligned: r_hi | r_lo | a0_hi | a0_lo
    float* mn = reinterpret_cast<float*>(
        smem_raw + (((smem_u32(raw + round_up(raw_floats(F), 4)) + 1023u) & ~1023u) - smem_u32(smem_raw)));
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int slot = warp < kBSlots * 4 ? (warp >> 2) : kBSlots;
    const int wtid = tid & 127;
    const int j = blockIdx.y, l = a.layer;
    const int N = a.N, B = a.B;
    const bool train = a.train != 0;
    constexpr int CT = kBSlots * 128;
    const bool is_compute = slot < kBSlots;

    LayerSrc src;
    src.params = a.params + (size_t)(j * L + l) * a.d.rec_stride;
    src.bn = a.bnbuf + (size_t)(j * L + l) * 8 * F;
    src.film = nullptr;
    src.mom = a.mom_in ? a.mom_in + j * GWTF_MOM_STRIDE : nullptr;
    src.sum1 = a.sum1 ? a.sum1 + (size_t)j * 4 * F : nullptr;
    src.n_total = a.n_total;

    if (warp == kBSlots * 4) tmem_alloc(&S.tmem_base, 512);
    if (tid == 0) {
        mbar_init(&S.bar_tma, 1);
        for (int s = 0; s < kBSlots; ++s) { mbar_init(&S.req[s], 128); mbar_init(&S.done[s], 1); }
        mbar_init(&S.buf_free, 1);
        mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_trigger();                    // only after this CTA owns its tensor-memory columns
    if (tid == 0) issue_layer_copy(raw, src, F, !train, false, &S.bar_tma);
    mbar_wait(&S.bar_tma, 0u);
    pdl_wait();                       // phase 0 of this layer (its sums, dobuf, gbuf) is complete from here on
    stage_vectors<FPN, true>(S.W, &S.WB, raw, src, F, a.d.warp_mask[l], train, false,
                             train ? a.bsum + (size_t)j * 8 * F : nullptr, tid, kBwdThreads);
    __syncthreads();

    const unsigned wm = a.d.warp_mask[l];
    const int w = popc3(wm), k = 3 - w;
    const NetOffsets o = net_offsets(F, w);
    int keepd[2];
    keepd[0] = !(wm & 1u) ? 0 : (!(wm & 2u) ? 1 : 2);                        // first / second kept dimension
    keepd[1] = k == 2 ? (!(wm & 4u) ? 2 : 1) : keepd[0];
    float* dpr = a.dparams + (size_t)(j * L + l) * a.d.rec_stride;
    double* bs = a.bsum + (size_t)j * 8 * F;
    const uint32_t tbase = S.tmem_base;

    const int tps = (N + 127) / 128;
    const int total_tiles = B * tps;
    const int per_cta = (total_tiles + gridDim.x - 1) / gridDim.x;
    const int t_begin = min(blockIdx.x * per_cta, total_tiles), t_end = min(t_begin + per_cta, total_tiles);
    const int my_tiles = t_end - t_begin;

#pragma unroll 1
    for (int net = 0; net < 2; ++net) {
        // ---- operands of this net that do not depend on the shape
        __syncthreads();              // everybody is done with the previous net's operands
        for (int i = tid; i < FPN * 8; i += kBwdThreads) {                     // B0 (as stage_b0)
            int e, kk;
            TcOperand<FPN, 8>::coords(i, e, kk);
            float v = 0.f;
            if (e < F) { const float4 q = S.W.q0[net][e]; v = kk == 0 ? q.x : (kk == 1 ? q.y : (kk == 2 ? q.z : (kk == 3 ? q.w : 0.f))); }
            else if (e == F && kk == 3) v = 1.f;
            S.B0.set(e, kk, v);
        }
        for (int i = tid; i < FPN * FPK; i += kBwdThreads) {                   // W1T[e][f] = W1[f][e]
            int e, f;
            TcOperand<FPN, FPK>::coords(i, e, f);
            S.W1T.set(e, f, (e < F && f < F) ? raw[net * o.stride + o.W1 + f * F + e] : 0.f);
        }
        for (int i = tid; i < 16 * FPK; i += kBwdThreads) {                    // Q0[d][e] = q0[e].d
            int d, e;
            TcOperand<16, FPK>::coords(i, d, e);
            float v = 0.f;
            if (d < 3 && e < F) { const float4 q = S.W.q0[net][e]; v = d == 0 ? q.x : (d == 1 ? q.y : q.z); }
            S.Q0.set(d, e, v);
        }
        for (int i = tid; i < 4 * 32; i += kBwdThreads) S.red[i] = 0.f;
        if (is_compute) {                                                      // zero this slot's dW1 accumulator
            float z[FPK];
#pragma unroll
            for (int i = 0; i < FPK; ++i) z[i] = 0.f;
            tmem_st<FPK>(tbase + C::G + slot * FPK + ((uint32_t)((warp & 3) * 32) << 16), z);
            tmem_wait_st();
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();

        if (!is_compute) {
            // ================= MMA issuer of slot s =================
            const int s = warp - kBSlots * 4;
            if (elect_one()) {
                uint32_t req_phase = (uint32_t)(net & 1);        // 4 requests per tile and one drain per net
                const uint32_t ts = tbase + s * C::SLOT;
                const uint32_t tg = tbase + C::G + s * FPK;
                SlotTurns turns{net * my_tiles};
                RoundIter rounds(t_begin, t_end, tps, nullptr, B, kBSlots);
                int rbase, rcount, rb;
#pragma unroll 1
                while (rounds.next(rbase, rcount, rb)) {
                    int off, my_seq;
                    if (!turns.take(s, rcount, off, my_seq)) continue;
                    mbar_wait(&S.req[s], req_phase); req_phase ^= 1u; tc_fence_after();
                    issue_ts<8, FPN>(ts + C::D, ts + C::Ahi, ts + C::Alo, S.B0.hi, S.B0.lo);
                    issue_ts<8, FPN>(ts + C::P, ts + C::Ahi, ts + C::Alo, S.PW.hi, S.PW.lo);
                    tc_commit(&S.done[s]);
                    mbar_wait(&S.req[s], req_phase); req_phase ^= 1u; tc_fence_after();
                    issue_ts<FPK, FPN>(ts + C::D, ts + C::Ahi, ts + C::Alo, S.B1.hi, S.B1.lo);
                    tc_commit(&S.done[s]);
                    mbar_wait(&S.req[s], req_phase); req_phase ^= 1u; tc_fence_after();
                    issue_ts<FPK, FPN>(ts + C::Ahi, ts + C::D, ts + C::P, S.W1T.hi, S.W1T.lo);     // r lives in D | P
                    tc_commit(&S.done[s]);
                    issue_point_contraction<FPK>(tg, mn);
                    tc_commit(&S.buf_free);                                     // contraction my_seq has consumed the buffer
                    mbar_wait(&S.req[s], req_phase); req_phase ^= 1u; tc_fence_after();
                    issue_ts<FPK, 16>(ts + C::Ahi, ts + C::D, ts + C::P, S.Q0.hi, S.Q0.lo);        // dy0 lives in D | P
                    tc_commit(&S.done[s]);
                }
                // drain: everything this thread issued (its dW1 chain included) has completed
                mbar_wait(&S.req[s], req_phase); tc_fence_after();
                tc_commit(&S.done[s]);
            }
            __syncwarp();
        } else {
            // ================= compute warpgroups =================
            const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
            const uint32_t trow = tbase + slot * C::SLOT + lane_off;
            uint64_t* req = &S.req[slot];
            uint64_t* done = &S.done[slot];
            uint32_t done_phase = (uint32_t)(net & 1);           // 4 requests per tile and one drain per net
            auto request = [&]() { tc_fence_before(); mbar_arrive(req); };
            auto wait_done = [&]() { mbar_wait(done, done_phase); done_phase ^= 1u; tc_fence_after(); };
            // column sums of dy0 (1 | xa | xb): acc0 = channel `lane`, acc1 = channel 32 + (lane & 7)
            float acc0[3] = {0.f, 0.f, 0.f}, acc1[3] = {0.f, 0.f, 0.f};
            float* coltile = S.col[warp];
            int cur_b = -1;
            SlotTurns turns{net * my_tiles};       // sequence numbers of this CTA's point contractions (buffer turns)
            GWTF_CLK_INIT(blockIdx.x == 0 && blockIdx.y == 0 && tid == 0)
            RoundIter it(t_begin, t_end, tps, nullptr, B, kBSlots);
            FilmAhead<FPN> film;
            film.b = -1; film.s = 0.f; film.t = 0.f;
            if (t_begin < t_end) film.fetch(a.film, it.b, B, K, j, L, l, F, tid);
            // the global loads of this slot's next tile (x, dO, incoming gradient) are issued while the current tile is
            // processed: within a shape that is tile t + 2.  The values carry the tile they belong to; at a shape boundary
            // (or if the guess was wrong) the tile loads its own.
            float pf[9];
            int pf_tile = -1;
            auto fetch = [&](int tile, int bb, int shape_begin) {
                pf_tile = tile;
#pragma unroll
                for (int i = 0; i < 9; ++i) pf[i] = 0.f;
                const int pn = (tile - shape_begin) * 128 + wtid;
                if (pn >= N) return;
                const float* pxin = a.xin_shared ? a.xin + (size_t)bb * 3 * N : a.xin + ((size_t)j * B + bb) * 3 * N;
                const float* pdob = a.dobuf + ((size_t)j * B + bb) * 6 * N + (size_t)net * 3 * N;
                const float* pg = a.gbuf + ((size_t)j * B + bb) * 3 * N;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    pf[d] = pxin[(size_t)d * N + pn];
                    pf[3 + d] = pdob[(size_t)d * N + pn];
                    pf[6 + d] = pg[(size_t)d * N + pn];
                }
            };
            int base, count, b;
            while (it.next(base, count, b)) {
                if (b != cur_b) {
                    // new shape: every warpgroup has drained the MMAs that read the per-shape operands
                    bwd_compute_barrier();
                    if (film.b != b) film.fetch(a.film, b, B, K, j, L, l, F, tid);
                    film.template stage<true>(S.W, &S.WB, F, tid);
                    film.fetch(a.film, b + 1, B, K, j, L, l, F, tid);  // in flight until the next shape boundary
                    bwd_compute_barrier();
                    for (int i = tid; i < FPN * FPK; i += CT) {                // B1 (as stage_b1 with the FiLM fold)
                        int f, e;
                        TcOperand<FPN, FPK>::coords(i, f, e);
                        float v = 0.f;
                        if (f < F) {
                            if (e < F) v = S.W.st[net][f].x * raw[net * o.stride + o.W1 + f * F + e];
                            else if (e == F) v = S.W.st[net][f].y;
                        } else if (f == F && e == F) v = 1.f;
                        S.B1.set(f, e, v);
                    }
                    for (int i = tid; i < FPN * 8; i += CT) {                  // PW[f][4+d] = (s/sigma1) w2[f].d
                        int f, kk;
                        TcOperand<FPN, 8>::coords(i, f, kk);
                        float v = 0.f;
                        if (f < F && kk >= 4 && kk < 7) {
                            const float4 w2 = S.W.w2[net][f];
                            v = S.W.st[net][f].x * (kk == 4 ? w2.x : (kk == 5 ? w2.y : w2.z));
                        }
                        S.PW.set(f, kk, v);
                    }
                    for (int f = tid; f < FPN; f += CT) {                      // dh1 = [y1>0] P + beta y1 + gamma
                        float2 v = make_float2(0.f, 0.f);
                        if (f < F && train) {
                            const float s = S.WB.sg[net][f].x;
                            const float tt = S.W.st[net][f].y + s * S.W.mi1[net][f].x;   // t = st.y + s mean1 istd1
                            const float2 mi = S.W.mi1[net][f], ab = S.WB.ab1[net][f];
                            v = make_float2(-mi.y * ab.y / s, mi.y * (ab.y * tt / s - ab.x));
                        }
                        S.bg[f] = v;
                    }
                    fence_proxy_async();
                    bwd_compute_barrier();
                    cur_b = b;
                    GWTF_CLK(12)
                }
                int off, my_seq;
                if (!turns.take(slot, count, off, my_seq)) continue;
                const int t = base + off;
                if (pf_tile != t) fetch(t, b, it.shape_begin());
                const int n = (t - it.shape_begin()) * 128 + wtid;
                const bool valid = n < N;
                const size_t sb = ((size_t)j * B + b) * 3 * N;
                float x[3], dO[3], gold[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) { x[d] = pf[d]; dO[d] = pf[3 + d]; gold[d] = pf[6 + d]; }
                {   // the (x, 1 | dO) operand row of this point, in tensor memory where a0 will go once y0 has been read:
                    // no shared-memory write, hence no proxy fence (MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, ~1000 cycles here)
                    float h[8], lo[8];
#pragma unroll
                    for (int d = 0; d < 3; ++d) { split_tf32(x[d], h[d], lo[d]); split_tf32(dO[d], h[4 + d], lo[4 + d]); }
                    h[3] = 1.0f; lo[3] = 0.f; h[7] = 0.f; lo[7] = 0.f;
                    tmem_st8(trow + C::Ahi, h);
                    tmem_st8(trow + C::Alo, lo);
                    tmem_wait_st();
                }
                GWTF_CLK(0)
                request();                                              // -> batch A: y0, P
                wait_done();
                GWTF_CLK(1)
                uint32_t m0lo = 0u, m0hi = 0u;                          // [y0 > 0], channel c -> bit c
                {
                    float y[FPK];
                    tmem_ld<FPK>(trow + C::D, y);                       // one round trip for the whole row
                    tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < FPK; c += 8) {
                        float hi[8], lo[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const bool pos = y[c + i] > 0.f;
                            if (c + i < 32) m0lo |= pos ? (1u << (c + i)) : 0u; else m0hi |= pos ? (1u << (c + i - 32)) : 0u;
                            split_tf32(fmaxf(y[c + i], 0.f), hi[i], lo[i]);
                        }
                        tmem_st8(trow + C::Ahi + c, hi);
                        tmem_st8(trow + C::Alo + c, lo);
                    }
                }
                tmem_wait_st();
                GWTF_CLK(2)
                request();                                              // -> batch B: y1
                wait_done();
                GWTF_CLK(3)
                // r = dh1 = [y1 > 0] P + beta y1 + gamma, written over the y1 / P columns it was computed from (hi -> D,
                // lo -> P) for the da0 MMA, and together with a0 into the shared point-contraction operands.  Sixteen
                // channels per tensor-memory round trip; the operand-buffer turn (contraction my_seq - 1 consumed) is
                // only needed from the first shared-memory store on.
                bool have_turn = false;
#pragma unroll
                for (int c0 = 0; c0 < FPK; c0 += 16) {
                    constexpr int kMax = 16;
                    const int nch = FPK - c0 < kMax ? FPK - c0 : kMax;                  // 16, 16, 8 (compile time after unrolling)
                    float y1[kMax], p8[kMax];
                    if (nch == 16) { tmem_ld16(trow + C::D + c0, y1); tmem_ld16(trow + C::P + c0, p8); }
                    else { tmem_ld8(trow + C::D + c0, y1); tmem_ld8(trow + C::P + c0, p8); }
                    tmem_wait_ld();
                    float rh[kMax], rl[kMax];
#pragma unroll
                    for (int i = 0; i < kMax; i += 2) {
                        if (i < nch) {
                            const float4 g2 = *reinterpret_cast<const float4*>(&S.bg[c0 + i]);  // (beta, gamma) x 2
                            float r0 = fmaf(g2.x, y1[i], g2.y) + (y1[i] > 0.f ? p8[i] : 0.f);
                            float r1 = fmaf(g2.z, y1[i + 1], g2.w) + (y1[i + 1] > 0.f ? p8[i + 1] : 0.f);
                            if (!valid) { r0 = 0.f; r1 = 0.f; }
                            split_tf32(r0, rh[i], rl[i]);
                            split_tf32(r1, rh[i + 1], rl[i + 1]);
                        }
                    }
                    float ah[kMax], al[kMax];                           // a0 of these channels: in flight during the r stores
                    if (nch == 16) { tmem_ld16(trow + C::Ahi + c0, ah); tmem_ld16(trow + C::Alo + c0, al); }
                    else { tmem_ld8(trow + C::Ahi + c0, ah); tmem_ld8(trow + C::Alo + c0, al); }
#pragma unroll
                    for (int c = 0; c < kMax; c += 8) {
                        if (c < nch) {
                            tmem_st8(trow + C::D + c0 + c, rh + c);
                            tmem_st8(trow + C::P + c0 + c, rl + c);
                        }
                    }
                    if (!have_turn) {
                        GWTF_CLK(4)
                        // (contraction my_seq - 1 ran in the other slot -- SlotTurns -- and this slot's own my_seq - 2 is
                        // complete, so the barrier is at most one phase away)
                        if (my_seq > 0) mbar_wait(&S.buf_free, (uint32_t)((my_seq - 1) & 1));
                        have_turn = true;
                        GWTF_CLK(5)
                    }
#pragma unroll
                    for (int c = 0; c < kMax; c += 8) {
                        if (c < nch) {
                            float t8[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) t8[i] = rh[c + i];
                            store_mn_chunk(mn, 0, (c0 + c) >> 3, wtid, t8);
#pragma unroll
                            for (int i = 0; i < 8; ++i) t8[i] = rl[c + i];
                            store_mn_chunk(mn, 1, (c0 + c) >> 3, wtid, t8);
                        }
                    }
                    tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < kMax; c += 8) {
                        if (c < nch) {
                            float t8[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) t8[i] = ah[c + i];
                            store_mn_chunk(mn, 2, (c0 + c) >> 3, wtid, t8);
#pragma unroll
                            for (int i = 0; i < 8; ++i) t8[i] = al[c + i];
                            store_mn_chunk(mn, 3, (c0 + c) >> 3, wtid, t8);
                        }
                    }
                }
                tmem_wait_st();
                tc_fence_before();
                fence_proxy_async();
                GWTF_CLK(6)
                request();                                              // -> batch C: da0, then dW1 += r^T a0
                // (the proxy fence above is a MEMBAR.ALL.CTA in SASS and would wait for these loads: issue them after it,
                // with the rest of the tile to land in)
                if (t + 2 < min(it.shape_begin() + tps, t_end)) fetch(t + 2, b, it.shape_begin());
                wait_done();                                            // (da0 only: dW1 keeps running)
                GWTF_CLK(7)
                {
                    const float xa = keepd[0] == 0 ? x[0] : (keepd[0] == 1 ? x[1] : x[2]);
                    const float xb = k == 2 ? (keepd[1] == 0 ? x[0] : (keepd[1] == 1 ? x[1] : x[2])) : 0.f;
                    // dy0 = [y0 > 0] da0 -> operand (hi -> D, lo -> P) and, raw, into row `lane` of the warp's column tile
                    float dy[FPK];
                    tmem_ld<FPK>(trow + C::Ahi, dy);
                    tmem_wait_ld();
                    __syncwarp();                                       // the previous tile's column readers are done
                    float4* crow = reinterpret_cast<float4*>(coltile + lane * kColPitch);
#pragma unroll
                    for (int c = 0; c < FPK; c += 8) {
                        float hi[8], lo[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const bool pos = (c + i < 32) ? ((m0lo >> (c + i)) & 1u) : ((m0hi >> (c + i - 32)) & 1u);
                            dy[c + i] = pos ? dy[c + i] : 0.f;
                            split_tf32(dy[c + i], hi[i], lo[i]);
                        }
                        tmem_st8(trow + C::D + c, hi);
                        tmem_st8(trow + C::P + c, lo);
                        crow[c / 4] = make_float4(dy[c], dy[c + 1], dy[c + 2], dy[c + 3]);
                        crow[c / 4 + 1] = make_float4(dy[c + 4], dy[c + 5], dy[c + 6], dy[c + 7]);
                    }
                    *reinterpret_cast<float2*>(coltile + 32 * kColPitch + 2 * lane) = make_float2(xa, xb);
                    tmem_wait_st();
                    GWTF_CLK(8)
                    request();                                          // -> batch D: du
                    // sums of dy0 (1 | xa | xb) over the warp's 32 points: lane L sums column L
                    __syncwarp();
                    const float* wts = coltile + 32 * kColPitch;
#pragma unroll 8
                    for (int pp = 0; pp < 32; ++pp) {
                        const float v = lane < FPK ? coltile[pp * kColPitch + lane] : 0.f;
                        const float2 x2 = *reinterpret_cast<const float2*>(wts + 2 * pp);
                        acc0[0] += v;
                        acc0[1] = fmaf(v, x2.x, acc0[1]);
                        acc0[2] = fmaf(v, x2.y, acc0[2]);
                    }
                    if (FPK > 32) {
                        float t0 = 0.f, t1 = 0.f, t2 = 0.f;
                        const int ch = 32 + (lane & 7), grp = lane >> 3;
#pragma unroll
                        for (int pp = 0; pp < 8; ++pp) {
                            const float v = ch < FPK ? coltile[col_tail_point(grp, pp) * kColPitch + ch] : 0.f;
                            const float2 x2 = *reinterpret_cast<const float2*>(wts + 2 * col_tail_point(grp, pp));
                            t0 += v;
                            t1 = fmaf(v, x2.x, t1);
                            t2 = fmaf(v, x2.y, t2);
                        }
                        t0 += __shfl_xor_sync(0xffffffffu, t0, 8);  t0 += __shfl_xor_sync(0xffffffffu, t0, 16);
                        t1 += __shfl_xor_sync(0xffffffffu, t1, 8);  t1 += __shfl_xor_sync(0xffffffffu, t1, 16);
                        t2 += __shfl_xor_sync(0xffffffffu, t2, 8);  t2 += __shfl_xor_sync(0xffffffffu, t2, 16);
                        acc1[0] += t0; acc1[1] += t1; acc1[2] += t2;
                    }
                }
                GWTF_CLK(9)
                wait_done();
                GWTF_CLK(10)
                {
                    float du[8];
                    tmem_ld8(trow + C::Ahi, du);
                    tmem_wait_ld();
                    if (valid) {
#pragma unroll
                        for (int d = 0; d < 3; ++d) a.gbuf[sb + (size_t)d * N + n] = gold[d] + du[d];
                    }
                }
                GWTF_CLK(11)
            }
            // ---- drain this slot's MMAs (the dW1 chain of the last tile included)
            request();
            wait_done();
            GWTF_CLK(13)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                if (lane < FPK) atomicAdd(&S.red[q * FPK + lane], acc0[q]);
                if (lane < 8 && 32 + lane < FPK) atomicAdd(&S.red[q * FPK + 32 + lane], acc1[q]);
            }
        }
        __syncthreads();
        // ---- flush the per-channel sums of this net: dbeta0 = S1, dW0 raw sums, dgamma0 = r0 . (Sx, S1)
        for (int e = tid; e < F; e += kBwdThreads) {
            const float s1 = S.red[e], sa = S.red[FPK + e], sbv = k == 2 ? S.red[2 * FPK + e] : 0.f;
            const float4 rv = S.WB.r0[net][e];
            const float ra = keepd[0] == 0 ? rv.x : (keepd[0] == 1 ? rv.y : rv.z);
            const float rb = k == 2 ? (keepd[1] == 0 ? rv.x : (keepd[1] == 1 ? rv.y : rv.z)) : 0.f;
            const float dg = fmaf(ra, sa, fmaf(rb, sbv, rv.w * s1));
            atomicAdd(&dpr[net * o.stride + o.g0 + e], dg);
            atomicAdd(&bs[(net * 4 + 3) * F + e], (double)dg);
            atomicAdd(&dpr[net * o.stride + o.b0 + e], s1);
            atomicAdd(&bs[(net * 4 + 2) * F + e], (double)s1);
            atomicAdd(&dpr[net * o.stride + o.W0 + e * k + 0], sa);
            if (k == 2) atomicAdd(&dpr[net * o.stride + o.W0 + e * k + 1], sbv);
        }
        // ---- flush dW1: the M = 64 accumulator keeps row f in lane 32 (f / 16) + f % 16
        if (is_compute) {
            const int f = (warp & 3) * 16 + (wtid & 31);
            float g[FPK];
            tmem_ld<FPK>(tbase + C::G + slot * FPK + ((uint32_t)((warp & 3) * 32) << 16), g);
            tmem_wait_ld();
            if ((wtid & 31) < 16 && f < F) {
#pragma unroll
                for (int e = 0; e < FPK; ++e)
                    if (e < F) atomicAdd(&dpr[net * o.stride + o.W1 + f * F + e], g[e]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kBSlots * 4) tmem_dealloc(tbase, 512);
}

// =============================================================================================
// Phase 0 (k_bwd_layer_tc<.., 0>): d(o_mu, o_lv) of every point, FiLM / sd2 gradients, sd1_bn backward sums.
// Per tile and net (logvar net first: its head fixes sigma, hence dO of both nets):
//   batch A   y0 = X B0^T  ->  a0 = relu(y0)                      batch B   y1 = a0 B1^T  (folded)
//   logvar net only:  a1 = relu(y1) -> batch C   o = a1 B2^T (N = 16)  -> softsign / sigma / dO, written to dobuf, gbuf
//   sums over points of dO_d a1 and dO_d [y1 > 0] for the warped dims d (column sums through the warp's tile, per-shape accumulators):
//     dW2 = sum dO a1;   dt = sum_d W2_d sum dO_d [y1>0];   ds = (sum_d W2_d sum dO_d a1 - t dt) / s     (a1 = [y1>0] y1)
//   so the CUDA cores never form da1 = W2^T dO per point and never touch a per-channel constant in the tile loop.
// =============================================================================================
constexpr int kDSlots = 4;
constexpr int kBwdDThreads = kDSlots * 128 + kDSlots * 32;

__device__ __forceinline__ void bwd_d_barrier() { asm volatile("bar.sync 2, %0;" ::"n"(kDSlots * 128) : "memory"); }

template <int FPK, int FPN>
struct TcBwdDSmem {
    LayerT<FPN> W;
    LayerWB<FPN> WB;
    TcLayerOps<FPK, FPN> ops;      // B0 | B1 (folded, per shape) | B2 of the current net
    float x_hi[kDSlots][128 * 8], x_lo[kDSlots][128 * 8];
    float corr[12];
    float red[4 * FPK];             // [a1 dA | a1 dB | m dA | m dB] per channel (m = [y1 > 0])
    alignas(16) float col[kDSlots * 4][kColTile];
    uint64_t bar_tma, req[kDSlots], done[kDSlots];
    uint32_t tmem_base;
};

template <int FPK, int FPN>
__device__ __forceinline__ void bwd_tc_phase0(const BwdArgs& a, unsigned char* smem_raw) {
    using SM = TcBwdDSmem<FPK, FPN>;
    using C = TcCols<FPK, FPN>;
    SM& S = *reinterpret_cast<SM*>(smem_raw);
    float* raw = reinterpret_cast<float*>(smem_raw + round_up((int)sizeof(SM), 16));
    const int F = a.d.n_features, K = a.d.n_components, L = a.d.n_layers;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int slot = warp < kDSlots * 4 ? (warp >> 2) : kDSlots;
    const int wtid = tid & 127;
    const int j = blockIdx.y, l = a.layer;
    const int N = a.N, B = a.B;
    const bool train = a.train != 0;
    constexpr int CT = kDSlots * 128;
    const bool is_compute = slot < kDSlots;

    LayerSrc src;
    src.params = a.params + (size_t)(j * L + l) * a.d.rec_stride;
    src.bn = a.bnbuf + (size_t)(j * L + l) * 8 * F;
    src.film = nullptr;
    src.mom = a.mom_in ? a.mom_in + j * GWTF_MOM_STRIDE : nullptr;
    src.sum1 = a.sum1 ? a.sum1 + (size_t)j * 4 * F : nullptr;
    src.n_total = a.n_total;

    if (warp == kDSlots * 4) tmem_alloc(&S.tmem_base, 512);
    if (tid == 0) {
        mbar_init(&S.bar_tma, 1);
        for (int s = 0; s < kDSlots; ++s) { mbar_init(&S.req[s], 128); mbar_init(&S.done[s], 1); }
        mbar_fence_init();
    }
    for (int i = tid; i < kDSlots * 128 * 8; i += kBwdDThreads) { (&S.x_hi[0][0])[i] = 0.f; (&S.x_lo[0][0])[i] = 0.f; }
    if (tid < 12) S.corr[tid] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_trigger();
    if (tid == 0) issue_layer_copy(raw, src, F, !train, false, &S.bar_tma);
    mbar_wait(&S.bar_tma, 0u);
    pdl_wait();                       // the previous layer's phase 1 (its gbuf, bn0 sums) is complete from here on
    const bool correct = train && a.mom_prev != nullptr;
    if (correct)
        bn0_correction<FPN>(a.d, a.params, j, l - 1, a.mom_prev + j * GWTF_MOM_STRIDE, a.bsum_prev + (size_t)j * 8 * F,
                            a.n_total, S.corr, tid);
    stage_vectors<FPN, true>(S.W, &S.WB, raw, src, F, a.d.warp_mask[l], train, false, nullptr, tid, kBwdDThreads);
    __syncthreads();
    float Mc[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) Mc[i] = correct ? S.corr[i] : 0.f;

    const unsigned wm = a.d.warp_mask[l];
    const int w = popc3(wm);
    const NetOffsets o = net_offsets(F, w);
    int wd[2];                                                   // first / second warped dimension
    wd[0] = (wm & 1u) ? 0 : ((wm & 2u) ? 1 : 2);
    wd[1] = w == 2 ? ((wm & 4u) ? 2 : 1) : wd[0];
    float* dpr = a.dparams + (size_t)(j * L + l) * a.d.rec_stride;
    double* bs = a.bsum + (size_t)j * 8 * F;
    const uint32_t tbase = S.tmem_base;

    const int tps = (N + 127) / 128;
    const int total_tiles = B * tps;
    const int per_cta = (total_tiles + gridDim.x - 1) / gridDim.x;
    const int t_begin = min(blockIdx.x * per_cta, total_tiles), t_end = min(t_begin + per_cta, total_tiles);

#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int net = 1 - pass;                                // the logvar net first
        const int stages = net == 1 ? 3 : 2;                     // MMA batches per tile
        __syncthreads();
        stage_b0<FPK, FPN>(S.ops, S.W.q0[net], F, tid, kBwdDThreads);
        stage_b2<FPK, FPN>(S.ops, S.W.w2[net], S.W.b2[net], F, tid, kBwdDThreads);
        for (int i = tid; i < 4 * FPK; i += kBwdDThreads) S.red[i] = 0.f;
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();

        // how many requests each slot made in the first pass (3 per tile + drain): barrier parities go on from there
        auto tiles_of = [&](int s) {
            int n = 0;
            RoundIter it0(t_begin, t_end, tps, nullptr, B, kDSlots);
            int base, count, b;
            while (it0.next(base, count, b)) n += (s < count) ? 1 : 0;
            return n;
        };

        if (!is_compute) {
            const int s = warp - kDSlots * 4;
            if (elect_one()) {
                const int tiles_s = tiles_of(s);
                uint32_t req_phase = pass == 0 ? 0u : (uint32_t)((tiles_s * 3 + 1) & 1);
                const uint32_t ts = tbase + s * kTcCols;
#pragma unroll 1
                for (int t = 0; t < tiles_s; ++t) {
                    mbar_wait(&S.req[s], req_phase); req_phase ^= 1u; tc_fence_after();
                    issue_ss_k8<FPN>(ts + C::D, S.x_hi[s], S.x_lo[s], S.ops.B0.hi, S.ops.B0.lo);
                    tc_commit(&S.done[s]);
                    mbar_wait(&S.req[s], req_phase); req_phase ^= 1u; tc_fence_after();
                    issue_ts<FPK, FPN>(ts + C::D, ts + C::Ahi, ts + C::Alo, S.ops.B1.hi, S.ops.B1.lo);
                    tc_commit(&S.done[s]);
                    if (stages == 3) {
                        mbar_wait(&S.req[s], req_phase); req_phase ^= 1u; tc_fence_after();
                        issue_ts<FPK, 16>(ts + C::D, ts + C::Ahi, ts + C::Alo, S.ops.B2.hi, S.ops.B2.lo);
                        tc_commit(&S.done[s]);
                    }
                }
                mbar_wait(&S.req[s], req_phase); tc_fence_after();      // drain
                tc_commit(&S.done[s]);
            }
            __syncwarp();
        } else {
            const uint32_t trow = tbase + slot * kTcCols + ((uint32_t)((warp & 3) * 32) << 16);
            uint64_t* req = &S.req[slot];
            uint64_t* done = &S.done[slot];
            uint32_t done_phase = pass == 0 ? 0u : (uint32_t)((tiles_of(slot) * 3 + 1) & 1);
            auto request = [&]() { tc_fence_before(); mbar_arrive(req); };
            auto wait_done = [&]() { mbar_wait(done, done_phase); done_phase ^= 1u; tc_fence_after(); };
            // column sums of this warp's points: acc0 = channel `lane`, acc1 = channel 32 + (lane & 7)
            float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};      // a1 dA | a1 dB | m dA | m dB
            float* coltile = S.col[warp];

            auto flush_shape = [&](int b) {
                // per-shape sums -> block partials -> FiLM / sd2 gradients and the sd1_bn backward sums
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (lane < FPK) atomicAdd(&S.red[q * FPK + lane], acc0[q]);
                    if (lane < 8 && 32 + lane < FPK) atomicAdd(&S.red[q * FPK + 32 + lane], acc1[q]);
                    acc0[q] = 0.f; acc1[q] = 0.f;
                }
                bwd_d_barrier();
                float* dfl = a.dfilm + ((size_t)(b * K + j) * L + l) * 4 * F;
                for (int f = tid; f < F; f += CT) {
                    const float4 w2 = S.W.w2[net][f];
                    const float s = S.WB.sg[net][f].x;
                    const float tt = S.W.st[net][f].y + s * S.W.mi1[net][f].x;
                    float Asum = 0.f, dt = 0.f;
                    for (int q = 0; q < w; ++q) {
                        const float wq = wd[q] == 0 ? w2.x : (wd[q] == 1 ? w2.y : w2.z);
                        const float sa = S.red[q * FPK + f], sm = S.red[(2 + q) * FPK + f];
                        Asum = fmaf(wq, sa, Asum);
                        dt = fmaf(wq, sm, dt);
                        atomicAdd(&dpr[net * o.stride + o.W2 + q * F + f], sa);
                    }
                    const float ds = (Asum - tt * dt) / s;
                    atomicAdd(&dfl[net * 2 * F + f], ds);
                    atomicAdd(&dfl[net * 2 * F + F + f], dt);
                    if (train) {
                        atomicAdd(&bs[(net * 4 + 1) * F + f], (double)(ds * s));
                        atomicAdd(&bs[(net * 4 + 0) * F + f], (double)(dt * s));
                    }
                }
                if (tid < w) atomicAdd(&dpr[net * o.stride + o.b2 + tid], S.red[tid * FPK + F]);
                bwd_d_barrier();
                for (int i = tid; i < 4 * FPK; i += CT) S.red[i] = 0.f;
            };

            int cur_b = -1;
            GWTF_CLK_INIT(blockIdx.x == 0 && blockIdx.y == 0 && tid == 0)
            RoundIter it(t_begin, t_end, tps, nullptr, B, kDSlots);
            int base, count, b;
            while (it.next(base, count, b)) {
                if (b != cur_b) {
                    if (cur_b >= 0) flush_shape(cur_b);
                    bwd_d_barrier();
                    stage_film<FPN, true>(S.W, &S.WB, a.film + ((size_t)(b * K + j) * L + l) * 4 * F, F, tid, CT);
                    bwd_d_barrier();
                    stage_b1<FPK, FPN>(S.ops, raw + net * o.stride + o.W1, S.W.st[net], F, tid, CT);
                    fence_proxy_async();
                    bwd_d_barrier();
                    cur_b = b;
                    GWTF_CLK(24)
                }
                if (slot >= count) continue;
                const int t = base + slot;
                const int n = (t - it.shape_begin()) * 128 + wtid;
                const bool valid = n < N;
                const float* xin = a.xin_shared ? a.xin + (size_t)b * 3 * N : a.xin + ((size_t)j * B + b) * 3 * N;
                const size_t sb = ((size_t)j * B + b) * 3 * N;
                float* dob = a.dobuf + ((size_t)j * B + b) * 6 * N;
                float x[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) x[d] = valid ? xin[(size_t)d * N + n] : 0.f;
                float outv[3] = {0.f, 0.f, 0.f}, gc[3] = {0.f, 0.f, 0.f}, gsv[3] = {0.f, 0.f, 0.f}, dO[3] = {0.f, 0.f, 0.f};
                if (net == 1) {
                    if (valid) {
                        float g0[3];
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            outv[d] = a.xout[sb + (size_t)d * N + n];
                            g0[d] = a.gbuf[sb + (size_t)d * N + n];
                            gsv[d] = a.gs[sb + (size_t)d * N + n];
                        }
#pragma unroll
                        for (int d = 0; d < 3; ++d)        // lazy bn0 correction of the layer processed before this one
                            gc[d] = g0[d] - (Mc[d * 3] * outv[0] + Mc[d * 3 + 1] * outv[1] + Mc[d * 3 + 2] * outv[2]) + Mc[9 + d];
                    }
                } else if (valid) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) dO[d] = dob[(size_t)d * N + n];
                }
                write_x_operand(S.x_hi[slot], S.x_lo[slot], x, wtid);
                fence_proxy_async();
                GWTF_CLK(16)
                request();                                              // -> y0
                wait_done();
                GWTF_CLK(17)
                relu_to_operand<FPK, FPN>(trow);
                GWTF_CLK(18)
                request();                                              // -> y1
                wait_done();
                GWTF_CLK(19)
                float y1[FPK];
                tmem_ld<FPK>(trow + C::D, y1);
                tmem_wait_ld();
                if (net == 1) {
                    // head of the logvar net on the tensor core: a1 -> operand, o = a1 B2^T
#pragma unroll
                    for (int c = 0; c < FPK; c += 8) {
                        float hi[8], lo[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) split_tf32(fmaxf(y1[c + i], 0.f), hi[i], lo[i]);
                        tmem_st8(trow + C::Ahi + c, hi);
                        tmem_st8(trow + C::Alo + c, lo);
                    }
                    tmem_wait_st();
                    GWTF_CLK(20)
                    request();                                          // -> o
                    wait_done();
                    GWTF_CLK(21)
                    float ov[8];
                    tmem_ld8(trow + C::D, ov);
                    tmem_wait_ld();
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const float olv = ov[d];
                        const float lam = softsign(olv);
                        const float ex = expf(lam);
                        const float sig2 = GWTF_FLOW_EPS + ex;
                        const float sig = sqrtf(sig2);
                        const float gin = gc[d] / sig;
                        const float dlam = gsv[d] - gc[d] * outv[d] * ex / (2.0f * sig2);
                        const float den = 1.0f + fabsf(olv);
                        const float dov = valid ? dlam / (den * den) : 0.f;
                        if (valid) {
                            a.gbuf[sb + (size_t)d * N + n] = gin;
                            dob[(size_t)d * N + n] = -gin;
                            dob[(size_t)(3 + d) * N + n] = dov;
                        }
                        dO[d] = dov;
                    }
                }
                GWTF_CLK(22)
                // ---- sums over the warp's 32 points: rows of a1 into the warp's tile, columns summed by the reader lanes.
                // The constant-one channel F has a1 = 1, so its a1 dA / a1 dB sums ARE the sd2 bias gradients.
                const float dA = wd[0] == 0 ? dO[0] : (wd[0] == 1 ? dO[1] : dO[2]);
                const float dB = w == 2 ? (wd[1] == 0 ? dO[0] : (wd[1] == 1 ? dO[1] : dO[2])) : 0.f;
#pragma unroll
                for (int e = 0; e < FPK; ++e) y1[e] = fmaxf(y1[e], 0.f);
                __syncwarp();                                           // the previous tile's readers are done
                col_write_row<FPK>(coltile, lane, y1, dA, dB);
                __syncwarp();
                {
                    const float* wts = coltile + 32 * kColPitch;
#pragma unroll 8
                    for (int pp = 0; pp < 32; ++pp) {
                        const float v = lane < FPK ? coltile[pp * kColPitch + lane] : 0.f;
                        const float2 d2 = *reinterpret_cast<const float2*>(wts + 2 * pp);
                        acc0[0] = fmaf(v, d2.x, acc0[0]);
                        acc0[1] = fmaf(v, d2.y, acc0[1]);
                        if (v > 0.f) { acc0[2] += d2.x; acc0[3] += d2.y; }
                    }
                    if (FPK > 32) {
                        float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
                        const int ch = 32 + (lane & 7), grp = lane >> 3;
#pragma unroll
                        for (int pp = 0; pp < 8; ++pp) {
                            const float v = ch < FPK ? coltile[col_tail_point(grp, pp) * kColPitch + ch] : 0.f;
                            const float2 d2 = *reinterpret_cast<const float2*>(wts + 2 * col_tail_point(grp, pp));
                            t0 = fmaf(v, d2.x, t0);
                            t1 = fmaf(v, d2.y, t1);
                            if (v > 0.f) { t2 += d2.x; t3 += d2.y; }
                        }
                        t0 += __shfl_xor_sync(0xffffffffu, t0, 8);  t0 += __shfl_xor_sync(0xffffffffu, t0, 16);
                        t1 += __shfl_xor_sync(0xffffffffu, t1, 8);  t1 += __shfl_xor_sync(0xffffffffu, t1, 16);
                        t2 += __shfl_xor_sync(0xffffffffu, t2, 8);  t2 += __shfl_xor_sync(0xffffffffu, t2, 16);
                        t3 += __shfl_xor_sync(0xffffffffu, t3, 8);  t3 += __shfl_xor_sync(0xffffffffu, t3, 16);
                        acc1[0] += t0; acc1[1] += t1; acc1[2] += t2; acc1[3] += t3;
                    }
                }
                GWTF_CLK(23)
            }
            if (cur_b >= 0) flush_shape(cur_b);
            GWTF_CLK(24)
            request();                                                  // drain
            wait_done();
            GWTF_CLK(25)
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kDSlots * 4) tmem_dealloc(tbase, 512);
}

template <int FPK, int FPN>
__host__ __device__ constexpr size_t bwd_tc_d_smem(int F) {
    return round_up((int)sizeof(TcBwdDSmem<FPK, FPN>), 16) + (size_t)round_up(raw_floats(F), 4) * 4;
}

template <int FPK, int FPN>
__host__ __device__ constexpr size_t bwd_tc_smem(int F) {
    return round_up((int)sizeof(TcBwdESmem<FPK, FPN>), 16) + (size_t)round_up(raw_floats(F), 4) * 4 + 1024 +
           (size_t)kMnTotal * 4;
}

template <int FPK, int FPN, int PHASE>
__global__ void __launch_bounds__(PHASE == 0 ? kBwdDThreads : kBwdThreads, 1) k_bwd_layer_tc(const BwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    GWTF_CLK_ZERO()
    if constexpr (PHASE == 0) bwd_tc_phase0<FPK, FPN>(a, smem_raw);
    else bwd_tc_phase1<FPK, FPN>(a, smem_raw);
    exchange_tail(a.tail);
    GWTF_CLK_FLUSH()
}

}  // namespace gwtf
