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
//   A warpgroup = 128 threads = 128 TMEM lanes (thread t owns point t of the tile); per tile slot the TMEM columns
//   are [0,FPN) accumulator | [FPN,FPN+FPK) A hi | [FPN+FPK,FPN+2FPK) A lo.  The kernel that drives these
//   building blocks is the persistent warp-specialised one in gwtf_tc_persist.cuh.
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
    // element (row, k) stored at float offset i: a staging loop that walks i has its 32 lanes on 32 distinct banks (walking
    // rows x k instead puts 8 lanes on one bank: consecutive k-quads are 32 floats apart)
    static __device__ __forceinline__ void coords(int i, int& row, int& k) {
        const int blk = i >> 5, lane = i & 31;
        const int rg = blk / (KT / 4), kg = blk - rg * (KT / 4);
        row = rg * 8 + (lane >> 2);
        k = kg * 4 + (lane & 3);
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
        int e, k;
        TcOperand<FPN, 8>::coords(i, e, k);
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
        int f, e;
        TcOperand<FPN, FPK>::coords(i, f, e);
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
        int d, f;
        TcOperand<16, FPK>::coords(i, d, f);
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
// PASSES = 3: fp32-grade 3xTF32; PASSES = 1: single-pass TF32 (the hi operands only)
template <int KT, int N, int PASSES = 3>
__device__ __forceinline__ void issue_ts(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, const float* b_hi,
                                         const float* b_lo) {
    const uint32_t idesc = make_idesc_tf32(128, N, 0, 0);
    const uint64_t bh = make_smem_desc_kmajor(b_hi, KT), bl = make_smem_desc_kmajor(b_lo, KT);
    bool acc = false;
#pragma unroll
    for (int pass = 0; pass < PASSES; ++pass) {
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

template <int FPK, int FPN, int PASSES = 3>
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
    if (PASSES == 1) {
        // single-pass TF32: the tensor core reads the top 19 bits of the fp32 word; round to nearest first so the
        // error stays unbiased (relu + 2 integer ops per channel, one tcgen05.st instead of two)
#pragma unroll
        for (int c = 0; c < FPK; c += 8) {
            float hi[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                hi[i] = __uint_as_float((__float_as_uint(fmaxf(y[c + i], 0.f)) + 0x1000u) & 0xffffe000u);
            tmem_st8(trow + C::Ahi + c, hi);
        }
        tmem_wait_st();
        return;
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


// ---------------------------------------------------------------------------------------------
// Per-channel sums over the 32 points of a warp without the 31-shuffle reduce-scatter: every lane (= point) writes
// its row of channel values into the warp's private shared-memory tile (row pitch 44 floats: conflict-free 16-byte
// stores), then lane L walks DOWN column L (channels 0..31) over the 32 points, and the last 8 channels are done by
// 4 lanes each (8 points per lane, two xor-shuffles to combine).  ~6 instructions per (point, channel) pair on one
// lane instead of ~8 per value on all 32: 3-4x fewer issue slots for the 80-160 values of the backward phases.
// Accumulators stay in the reader lanes' registers across tiles.
// ---------------------------------------------------------------------------------------------
constexpr int kColPitch = 44;                       // floats per point row (40 channels + pad)
constexpr int kColTile = 32 * kColPitch + 64;       // floats per warp: the tile + 32 x (wA, wB) per-point weights

// writer: this lane's row (channels [0, 40)) and its two per-point weights
template <int NCH>
__device__ __forceinline__ void col_write_row(float* tile, int lane, const float (&v)[NCH], float wA, float wB) {
    static_assert(NCH % 4 == 0 && NCH <= 40, "row of at most 40 channels");
    float4* row = reinterpret_cast<float4*>(tile + lane * kColPitch);
#pragma unroll
    for (int c = 0; c < NCH / 4; ++c) row[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    *reinterpret_cast<float2*>(tile + 32 * kColPitch + 2 * lane) = make_float2(wA, wB);
}
// channel this lane accumulates in the two reader passes: pass 0 -> lane, pass 1 -> 32 + (lane & 7) (complete in
// every lane after the shuffles; lanes >= 8 hold copies).  In pass 1 lane group g = lane >> 3 reads the points below:
// at every step the four groups are 6 (or 2) rows apart, i.e. 8 / 24 banks, so the 32 lanes hit 32 distinct banks.
__device__ __forceinline__ int col_tail_point(int g, int pp) { return pp < 6 ? 6 * g + pp : 24 + 2 * g + (pp - 6); }

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

}  // namespace gwtf
