// Tensor-core (tcgen05) forward of one coupling layer, "all-GEMM" form.  Per 128-point tile and net
// the whole per-point MLP is three chained UMMAs with fp32-grade 3xTF32 splitting
// (X = hi + lo, A*B ~= Ahi*Bhi + Alo*Bhi + Ahi*Blo, fp32 accumulation in tensor memory):
//
//   MMA0  y0[128 x FPN] = X[128 x 8] * B0^T      X = (x0,x1,x2,1,0..): BN0-folded sd0 incl. bias
//   relu, split -> A1 (tcgen05.st, one TMEM lane per point)
//   MMA1  y1[128 x FPN] = A1[128 x FPK] * B1^T   B1 = diag(s/sigma1) W1 | shift column: sd1 + BN1 + FiLM
//   relu, split -> A2
//   MMA2  o [128 x 16 ] = A2[128 x FPK] * B2^T   B2 = sd2 weight rows scattered to xyz | bias column
//
// A constant-one channel (index F) carries the biases through the chain, so the CUDA cores only do
// relu + hi/lo split between the MMAs (3 instructions per channel, no shared-memory reads).
//   CTA = 128 threads = 128 TMEM lanes (thread t owns point t of the tile); one tile in flight per
//   CTA, 4 CTAs per SM (TMEM: [0,FPN) accumulator | [FPN,FPN+FPK) A hi | [FPN+FPK,FPN+2FPK) A lo).
#pragma once
#include "gwtf_fwd.cuh"
#include "gwtf_tc.cuh"
#include "gwtf_mma.cuh"

namespace gwtf {

#ifdef GWTF_TIMING
__device__ long long g_tc_cycles[16];
#define GWTF_T(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { long long c_ = clock64(); g_tc_cycles[i] += c_ - t_last_; t_last_ = c_; } } while (0)
#define GWTF_T0() long long t_last_ = clock64()
#define GWTF_RT(i) do {} while (0)
#else
#define GWTF_RT(i) do {} while (0)
#define GWTF_T(i) do {} while (0)
#define GWTF_T0() do {} while (0)
#endif

constexpr int kTcThreads = 128;
constexpr int kTcCols = 128;          // TMEM columns per CTA

template <int FPN>
struct LayerT {             // per-channel vectors (FPN entries per net), same member names as LayerW
    float4 q0[2][FPN];
    float2 st[2][FPN];
    float2 mi1[2][FPN];
    float4 w2[2][FPN];
    float4 b2[2];
};

template <int ROWS, int KT>
struct TcOperand {          // UMMA smem operand, K-major, 3xTF32 split
    float hi[ROWS * KT];
    float lo[ROWS * KT];
    __device__ __forceinline__ void set(int row, int k, float v) {
        float h, l;
        split_tf32(v, h, l);
        const int off = kmajor_offset(row, k, KT);
        hi[off] = h;
        lo[off] = l;
    }
};

template <int FPK, int FPN>
struct TcLayerOps {         // operands of one net
    TcOperand<FPN, 8> B0;       // [e][x0,x1,x2,1,0,0,0,0]
    TcOperand<FPN, FPK> B1;     // [f][e]
    TcOperand<16, FPK> B2;      // [d][f]
};

// B0 from the staged vectors (q0 = BN0-folded sd0 rows + bias); row F generates the constant one.
template <int FPK, int FPN>
__device__ __forceinline__ void stage_b0(TcLayerOps<FPK, FPN>& O, const float4* q0, int F, int tid, int nthreads) {
    for (int i = tid; i < FPN * 8; i += nthreads) {
        const int e = i >> 3, k = i & 7;
        float v = 0.f;
        if (e < F) {
            const float4 q = q0[e];
            v = k == 0 ? q.x : (k == 1 ? q.y : (k == 2 ? q.z : (k == 3 ? q.w : 0.f)));
        } else if (e == F && k == 3) v = 1.f;
        O.B0.set(e, k, v);
    }
}
// B1[f][e] = scale_f * W1[f][e], B1[f][F] = shift_f, B1[F][F] = 1.  st == nullptr: plain W1 (statistics pass)
template <int FPK, int FPN>
__device__ __forceinline__ void stage_b1(TcLayerOps<FPK, FPN>& O, const float* w1, const float2* st, int F, int tid,
                                         int nthreads) {
    for (int i = tid; i < FPN * FPK; i += nthreads) {
        const int f = i / FPK, e = i - f * FPK;
        float v = 0.f;
        if (f < F) {
            if (e < F) v = st ? st[f].x * w1[f * F + e] : w1[f * F + e];
            else if (e == F && st) v = st[f].y;
        } else if (f == F && e == F) v = 1.f;
        O.B1.set(f, e, v);
    }
}
// B2[d][f] = w2[f][d], B2[d][F] = b2[d]
template <int FPK, int FPN>
__device__ __forceinline__ void stage_b2(TcLayerOps<FPK, FPN>& O, const float4* w2, float4 b2, int F, int tid,
                                         int nthreads) {
    for (int i = tid; i < 16 * FPK; i += nthreads) {
        const int d = i / FPK, f = i - d * FPK;
        float v = 0.f;
        if (d < 3) {
            if (f < F) { const float4 w = w2[f]; v = d == 0 ? w.x : (d == 1 ? w.y : w.z); }
            else if (f == F) v = d == 0 ? b2.x : (d == 1 ? b2.y : b2.z);
        }
        O.B2.set(d, f, v);
    }
}

template <int FPK, int FPN>
struct TcCols {
    static_assert(FPN + 2 * FPK <= kTcCols, "tile does not fit its TMEM column budget");
    static constexpr uint32_t D = 0, Ahi = FPN, Alo = FPN + FPK;
};

// one thread: D[128 x N] = A * B^T with the 3xTF32 passes, A from TMEM (hi/lo column blocks)
template <int KT, int N>
__device__ __forceinline__ void issue_ts(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, const float* b_hi,
                                         const float* b_lo) {
    const uint32_t idesc = make_idesc_tf32(128, N, 0, 0);
    const uint64_t bh = make_smem_desc_kmajor(b_hi, KT), bl = make_smem_desc_kmajor(b_lo, KT);
    bool acc = false;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a = pass == 1 ? a_lo : a_hi;
        const uint64_t b = pass == 2 ? bl : bh;
#pragma unroll
        for (int s = 0; s < KT / 8; ++s) {
            mma_tf32_ts(d_tmem, a + 8 * s, b + (uint64_t)(16 * s), idesc, acc);
            acc = true;
        }
    }
}
// one thread: D[128 x N] = X * B0^T, X from shared memory (K = 8)
template <int N>
__device__ __forceinline__ void issue_ss_k8(uint32_t d_tmem, const float* x_hi, const float* x_lo, const float* b_hi,
                                            const float* b_lo) {
    const uint32_t idesc = make_idesc_tf32(128, N, 0, 0);
    const uint64_t xh = make_smem_desc_kmajor(x_hi, 8), xl = make_smem_desc_kmajor(x_lo, 8);
    const uint64_t bh = make_smem_desc_kmajor(b_hi, 8), bl = make_smem_desc_kmajor(b_lo, 8);
    mma_tf32_ss(d_tmem, xh, bh, idesc, false);
    mma_tf32_ss(d_tmem, xl, bh, idesc, true);
    mma_tf32_ss(d_tmem, xh, bl, idesc, true);
}

// all threads: relu the accumulator row of this thread and write it back as the next A operand
// (all FPK columns are loaded in one go so the TMEM read latency is paid once, not per chunk)
// address of the (m-tile, row group) of point (b, n) in the kept-activation buffer (gwtf_mma.cuh), or null:
// rows g and g+8 of an m-tile share it (their values interleave inside each 16-byte fragment element)
__device__ __forceinline__ float* keep_ptr(float* keep, int j, int net, int F, int B, int N, int b, int n) {
    if (!keep) return nullptr;
    const size_t p = (size_t)b * keep_npad(N) + n;
    return keep + keep_slab(F, B, N, j, net) + ((p >> 4) * (size_t)(((F + 7) / 8) * 32) + (p & 7) * 4) * 4;
}

template <int FPK, int FPN>
__device__ __forceinline__ void relu_to_operand(uint32_t trow, float* keepf = nullptr, int keep_nt = 0, bool valid = true) {
    using C = TcCols<FPK, FPN>;
    float y[FPK];
    tmem_ld<FPK>(trow + C::D, y);
    tmem_wait_ld();
    if (keepf) {
        // y1 kept for the backward pass in MMA C-fragment order: a 16-byte element = (row g: ch 2t, 2t+1 | row g+8:
        // ch 2t, 2t+1).  This thread is one row; lanes l and l^8 are rows g and g+8 of the same m-tile, so they swap
        // halves by shuffle and each writes whole elements (lane g: even channel pairs, lane g+8: odd pairs) --
        // a warp store then fills complete 32-byte sectors.  Rows beyond N hold zeros.
        const bool upper = (threadIdx.x & 8) != 0;
#pragma unroll
        for (int k = 0; k < FPK / 4; ++k) {
            const int nt = k >> 1, t0 = (k & 1) * 2;
            if (nt < keep_nt) {
                const float e0 = valid ? y[8 * nt + 2 * t0] : 0.f, e1 = valid ? y[8 * nt + 2 * t0 + 1] : 0.f;
                const float o0 = valid ? y[8 * nt + 2 * t0 + 2] : 0.f, o1 = valid ? y[8 * nt + 2 * t0 + 3] : 0.f;
                const float r0 = __shfl_xor_sync(0xffffffffu, upper ? e0 : o0, 8);
                const float r1 = __shfl_xor_sync(0xffffffffu, upper ? e1 : o1, 8);
                const float4 out = upper ? make_float4(r0, r1, o0, o1) : make_float4(e0, e1, r0, r1);
                *reinterpret_cast<float4*>(keepf + (nt * 32 + t0 + (upper ? 1 : 0)) * 4) = out;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < FPK; c += 8) {
        float hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split_tf32(fmaxf(y[c + i], 0.f), hi[i], lo[i]);
        tmem_st8(trow + C::Ahi + c, hi);
        tmem_st8(trow + C::Alo + c, lo);
    }
    tmem_wait_st();
}

// CTA-wide hand-off: everything written (TMEM / smem) by all threads is visible to the MMA issuer
__device__ __forceinline__ void tc_handoff() {
    tc_fence_before();
    __syncthreads();
}
__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t& phase) {
    mbar_wait(bar, phase);
    phase ^= 1u;
    tc_fence_after();
}

// write this thread's point as row `tid` of the X operand: (x0,x1,x2,1 | 0,0,0,0)
__device__ __forceinline__ void write_x_operand(float* x_hi, float* x_lo, const float (&x)[3], int tid) {
    float h[3], l[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) split_tf32(x[d], h[d], l[d]);
    const int off = kmajor_offset(tid, 0, 8);
    *reinterpret_cast<float4*>(x_hi + off) = make_float4(h[0], h[1], h[2], 1.0f);
    *reinterpret_cast<float4*>(x_lo + off) = make_float4(l[0], l[1], l[2], 0.0f);
}

template <int FPK, int FPN>
struct TcFwdSmem {
    LayerT<FPN> W;
    TcLayerOps<FPK, FPN> ops[2];
    float x_hi[128 * 8], x_lo[128 * 8];
    uint64_t bar_tma, bar_mma;
    uint32_t tmem_base;
    float red[2][2 * FPN + 32];
    double dred[16];
};

// MMA0 -> relu -> MMA1 for one net; on return y1 (phase 1) / h1 (phase 0) sits in TMEM columns [0,FPN)
template <int FPK, int FPN>
__device__ __forceinline__ void run_to_h1(TcFwdSmem<FPK, FPN>& S, int net, uint32_t tbase, uint32_t trow,
                                          uint32_t& phase, int tid) {
    using C = TcCols<FPK, FPN>;
    if (tid == 0) {
        tc_fence_after();
        issue_ss_k8<FPN>(tbase + C::D, S.x_hi, S.x_lo, S.ops[net].B0.hi, S.ops[net].B0.lo);
        tc_commit(&S.bar_mma);
    }
    tc_wait(&S.bar_mma, phase);
    GWTF_RT(11);
    relu_to_operand<FPK, FPN>(trow);
    tc_handoff();
    GWTF_RT(12);
    if (tid == 0) {
        tc_fence_after();
        issue_ts<FPK, FPN>(tbase + C::D, tbase + C::Ahi, tbase + C::Alo, S.ops[net].B1.hi, S.ops[net].B1.lo);
        tc_commit(&S.bar_mma);
    }
    tc_wait(&S.bar_mma, phase);
}

template <int FPK, int FPN, int PHASE>
__global__ void __launch_bounds__(kTcThreads) k_fwd_layer_tc(const LayerArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using SM = TcFwdSmem<FPK, FPN>;
    using C = TcCols<FPK, FPN>;
    SM& S = *reinterpret_cast<SM*>(smem_raw);
    float* raw = reinterpret_cast<float*>(smem_raw + round_up((int)sizeof(SM), 16));
    const int F = a.d.n_features, K = a.d.n_components, L = a.d.n_layers;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j = blockIdx.y, l = a.layer;
    const int N = a.N, B = a.B;
    const bool train = a.train != 0;
    constexpr int TT = kTcThreads;

    LayerSrc src;
    src.params = a.params + (size_t)(j * L + l) * a.d.rec_stride;
    src.bn = a.bnbuf + (size_t)(j * L + l) * 8 * F;
    src.film = nullptr;
    src.mom = a.mom_in ? a.mom_in + j * GWTF_MOM_STRIDE : nullptr;
    src.sum1 = a.sum1 ? a.sum1 + (size_t)j * 4 * F : nullptr;
    src.n_total = a.n_total;

    GWTF_T0();
    if (warp == 0) tmem_alloc(&S.tmem_base, kTcCols);
    if (tid == 0) { mbar_init(&S.bar_tma, 1); mbar_init(&S.bar_mma, 1); mbar_fence_init(); }
    for (int i = tid; i < 2 * (2 * FPN + 32); i += TT) (&S.red[0][0])[i] = 0.f;
    for (int i = tid; i < 128 * 8; i += TT) { S.x_hi[i] = 0.f; S.x_lo[i] = 0.f; }
    if (tid < 16) S.dred[tid] = 0.0;
    __syncthreads();
    if (tid == 0) issue_layer_copy(raw, src, F, !train, false, &S.bar_tma);
    mbar_wait(&S.bar_tma, 0u);
    GWTF_T(0);
    stage_vectors<FPN, false>(S.W, (LayerWB<FPN>*)nullptr, raw, src, F, a.d.warp_mask[l], train, PHASE == 0, nullptr,
                              tid, TT);
    __syncthreads();
    GWTF_T(1);
    const NetOffsets o = net_offsets(F, popc3(a.d.warp_mask[l]));
#pragma unroll
    for (int net = 0; net < 2; ++net) {
        stage_b0<FPK, FPN>(S.ops[net], S.W.q0[net], F, tid, TT);
        if (PHASE == 0) stage_b1<FPK, FPN>(S.ops[net], raw + net * o.stride + o.W1, nullptr, F, tid, TT);
        else stage_b2<FPK, FPN>(S.ops[net], S.W.w2[net], S.W.b2[net], F, tid, TT);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    GWTF_T(2);
    const uint32_t tbase = S.tmem_base;
    const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);
    uint32_t mma_phase = 0u;
    float macc = 0.f;
    int cur_b = -1;

    // tiles of 128 points, never straddling shapes; a CTA takes a CONTIGUOUS range so that it
    // re-stages the per-shape B1 operand only when it crosses a shape boundary
    const int total_tiles = B * a.tiles_per_shape;
    const int per_cta = (total_tiles + gridDim.x - 1) / gridDim.x;
    const int t_begin = blockIdx.x * per_cta, t_end = min(t_begin + per_cta, total_tiles);
    float mv[9];                                       // per-thread moment partials of the outputs
#pragma unroll
    for (int i = 0; i < 9; ++i) mv[i] = 0.f;
    // software prefetch: the next tile's coordinates (and running log-det sums) are loaded while the
    // current tile is in the MMA chain, so their global latency is off the critical path
    auto tile_coords = [&](int t, int& b, int& n, bool& valid) {
        b = t / a.tiles_per_shape;
        n = (t - b * a.tiles_per_shape) * TT + tid;
        valid = n < N;
    };
    auto load_x = [&](int t, float (&x)[3], float (&s3)[3]) {
        int b, n; bool valid;
        tile_coords(t, b, n, valid);
        const float* xin = a.xin_shared ? a.xin + (size_t)b * 3 * N : a.xin + ((size_t)j * B + b) * 3 * N;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            x[d] = valid ? xin[(size_t)d * N + n] : 0.f;
            s3[d] = (PHASE == 1 && a.ssum && valid) ? a.ssum[((size_t)j * B + b) * 3 * N + (size_t)d * N + n] : 0.f;
        }
    };
    float xn[3], sn[3];
    if (t_begin < t_end) load_x(t_begin, xn, sn);
    for (int t = t_begin; t < t_end; ++t) {
        int b, n; bool valid;
        tile_coords(t, b, n, valid);
        if (PHASE == 1 && b != cur_b) {
            __syncthreads();
            stage_film<FPN, false>(S.W, (LayerWB<FPN>*)nullptr, a.film + ((size_t)(b * K + j) * L + l) * 4 * F, F, tid,
                                   TT);
            __syncthreads();
#pragma unroll
            for (int net = 0; net < 2; ++net)
                stage_b1<FPK, FPN>(S.ops[net], raw + net * o.stride + o.W1, S.W.st[net], F, tid, TT);
            cur_b = b;
            GWTF_T(3);
        }
        float x[3], s3[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) { x[d] = xn[d]; s3[d] = sn[d]; }
        write_x_operand(S.x_hi, S.x_lo, x, tid);
        fence_proxy_async();
        tc_handoff();
        GWTF_T(4);
        if (t + 1 < t_end) load_x(t + 1, xn, sn);

        if (PHASE == 0) {
#pragma unroll 1
            for (int net = 0; net < 2; ++net) {
                run_to_h1<FPK, FPN>(S, net, tbase, trow, mma_phase, tid);
                GWTF_T(5);
                // per-channel sum h1, sum h1^2: 16 channels -> 32 values per warp reduce-scatter
                float h[FPN];
                tmem_ld<FPN>(trow + C::D, h);
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < FPN; c += 16) {
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float hv = valid ? h[c + i] : 0.f;
                        v[2 * i] = hv;
                        v[2 * i + 1] = hv * hv;
                    }
                    const float r = warp_reduce_scatter32(v, lane);
                    atomicAdd(&S.red[net][2 * c + lane], r);
                }
                tc_handoff();      // every thread is done with the accumulator before the next MMA0
                GWTF_T(6);
            }
        } else {
            float o3[2][3];
#pragma unroll 1
            for (int net = 0; net < 2; ++net) {
                run_to_h1<FPK, FPN>(S, net, tbase, trow, mma_phase, tid);
                GWTF_T(5);
                relu_to_operand<FPK, FPN>(trow, keep_ptr(a.y1out, j, net, F, B, N, b, n), (F + 7) / 8, valid);
                tc_handoff();
                GWTF_T(7);
                if (tid == 0) {
                    tc_fence_after();
                    issue_ts<FPK, 16>(tbase + C::D, tbase + C::Ahi, tbase + C::Alo, S.ops[net].B2.hi, S.ops[net].B2.lo);
                    tc_commit(&S.bar_mma);
                }
                tc_wait(&S.bar_mma, mma_phase);
                GWTF_T(8);
                float ov[8];
                tmem_ld8(trow + C::D, ov);
                tmem_wait_ld();
                o3[net][0] = ov[0]; o3[net][1] = ov[1]; o3[net][2] = ov[2];
                tc_handoff();
                GWTF_T(9);
            }
            float lam[3];
            if (a.direct) warp_point<true>(x, o3[0], o3[1], lam);
            else warp_point<false>(x, o3[0], o3[1], lam);
            if (valid) {
                const size_t base = ((size_t)j * B + b) * 3 * N + n;
#pragma unroll
                for (int d = 0; d < 3; ++d) a.xout[base + (size_t)d * N] = x[d];
                if (a.ld) a.ld[((size_t)j * B + b) * N + n] += lam[0] + lam[1] + lam[2];
                if (a.ssum)
#pragma unroll
                    for (int d = 0; d < 3; ++d) a.ssum[base + (size_t)d * N] = s3[d] + lam[d];
                if (a.trio) {
                    const size_t tb = (((size_t)j * 3) * B + b) * 3 * N + n;
                    const size_t ts = (size_t)B * 3 * N;
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        a.trio[tb + (size_t)d * N] = x[d];
                        a.trio[tb + ts + (size_t)d * N] = o3[0][d];
                        a.trio[tb + 2 * ts + (size_t)d * N] = lam[d];
                    }
                }
                mv[0] += x[0]; mv[1] += x[1]; mv[2] += x[2];
                mv[3] += x[0] * x[0]; mv[4] += x[0] * x[1]; mv[5] += x[0] * x[2];
                mv[6] += x[1] * x[1]; mv[7] += x[1] * x[2]; mv[8] += x[2] * x[2];
            }
            GWTF_T(10);
        }
    }
    if (PHASE == 1 && a.mom_out) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = i < 9 ? mv[i] : 0.f;
        macc = warp_reduce_scatter32(v, lane);
    }
    // ---- flush block partials
    __syncthreads();
    if (PHASE == 0) {
        for (int i = tid; i < 2 * 2 * FPN; i += TT) {
            const int net = i / (2 * FPN), idx = i - net * 2 * FPN, f = idx >> 1, which = idx & 1;
            if (f < F) atomicAdd(&a.sum1[((size_t)j * 2 + net) * 2 * F + which * F + f], (double)S.red[net][idx]);
        }
    } else if (a.mom_out) {
        if (lane < 9) atomicAdd(&S.dred[lane], (double)macc);
        __syncthreads();
        if (tid < 9) atomicAdd(&a.mom_out[j * GWTF_MOM_STRIDE + tid], S.dred[tid]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, kTcCols);
}

}  // namespace gwtf
