// Shared device helpers for the gwtf kernels: layer-record staging (TMA bulk copy -> BN/FiLM
// folding -> compute layout), the per-point coupling-net microkernel, warp reduce-scatter.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gwtf.h"

#define GWTF_BN_EPS 1e-5f
#define GWTF_FLOW_EPS 1e-6f       // flows.py:13 `eps` buffer
#define GWTF_LOG_2PI 1.8378770664093453f

namespace gwtf {

constexpr int kThreads = 256;

__host__ __device__ constexpr int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline int rec_stride_of(int F) { return round_up(2 * (F * F + 5 * F + 2), 4); }

// natural in-record offsets for one net of a layer with `w` warped dims (k = 3 - w kept)
struct NetOffsets { int W0, g0, b0, W1, W2, b2, stride; };
__host__ __device__ inline NetOffsets net_offsets(int F, int w) {
    int k = 3 - w;
    NetOffsets o;
    o.W0 = 0; o.g0 = F * k; o.b0 = o.g0 + F; o.W1 = o.b0 + F; o.W2 = o.W1 + F * F; o.b2 = o.W2 + w * F;
    o.stride = o.b2 + w;
    return o;
}

__device__ __forceinline__ int popc3(unsigned m) { return __popc(m & 7u); }

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D TMA bulk copy (cp.async.bulk) -- SASS: UBLKCP / SYNCS
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "GWTF_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra GWTF_DONE;\n"
        "bra GWTF_WAIT;\n"
        "GWTF_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// compute layout of one coupling layer in shared memory (both nets)
// ---------------------------------------------------------------------------------------------
template <int FP>
struct LayerW {
    float4 q0[2][FP];       // BN0-folded sd0: a0[e] = relu(q.x*x0 + q.y*x1 + q.z*x2 + q.w)
    float W1T[2][FP][FP];   // [net][e][f] = sd1.weight[f][e]  (zero padded)
    float2 st[2][FP];       // BN1+FiLM folded: a1[f] = relu(st.x*h1[f] + st.y)   (stage_film)
    float2 mi1[2][FP];      // n1[f] = h1[f]*mi.y - mi.x           (mi.x = mean1*istd1, mi.y = istd1)
    float4 w2[2][FP];       // sd2.weight rows scattered to xyz (zero rows on kept dims)
    float4 b2[2];           // sd2.bias scattered to xyz
};

template <int FP>
struct LayerWB {            // extras the backward phases need
    float4 r0[2][FP];       // hhat0[e] = r.x*x0 + r.y*x1 + r.z*x2 + r.w   (normalised sd0 output)
    float2 sg[2][FP];       // (FiLM scale s, bn0.weight gamma0)
    float2 ab1[2][FP];      // sd1_bn backward means (mean dn1, mean dn1*n1); zero in eval mode
};

// raw staging buffer: [params rec_stride][bn 8F][film 4F] floats
__host__ __device__ inline int raw_floats(int F) { return rec_stride_of(F) + 8 * F + 4 * F; }

struct LayerSrc {
    const float* params;    // record of (j,l)
    const float* bn;        // 8F running stats of (j,l)          (eval mode)
    const float* film;      // 4F: s_mu | t_mu | s_lv | t_lv of (b,j,l)   (may be null for phase 0)
    const double* mom;      // 16 doubles: input moments of (l,j)  (train mode)
    const double* sum1;     // [2][2][F] doubles                    (train mode, phase >= 1)
    double n_total;
};

// Issue the bulk copies of one layer record into `raw` (one elected thread).
__device__ __forceinline__ void issue_layer_copy(float* raw, const LayerSrc& s, int F, bool want_bn, bool want_film,
                                                 uint64_t* bar) {
    const int rs = rec_stride_of(F);
    uint32_t bytes = rs * 4u + (want_bn ? 32u * F : 0u) + (want_film ? 16u * F : 0u);
    mbar_expect_tx(bar, bytes);
    tma_bulk_g2s(raw, s.params, rs * 4u, bar);
    if (want_bn) tma_bulk_g2s(raw + rs, s.bn, 32u * F, bar);
    if (want_film) tma_bulk_g2s(raw + rs + 8 * F, s.film, 16u * F, bar);
}

// Batch statistics of sd0's output from the input moments (h0 = W0 x is linear):
// mean0 = W0 mu, var0 = W0 Cov W0^T (biased).  Returns via refs.
__device__ __forceinline__ void bn0_from_moments(const double* mom, double n, const float* W0row, int k,
                                                 const int* keepd, float& mean0, float& var0) {
    // mom: Sx(3), Sxx upper (00,01,02,11,12,22)
    double mu[3] = {mom[0] / n, mom[1] / n, mom[2] / n};
    double c[3][3];
    c[0][0] = mom[3] / n - mu[0] * mu[0];
    c[0][1] = c[1][0] = mom[4] / n - mu[0] * mu[1];
    c[0][2] = c[2][0] = mom[5] / n - mu[0] * mu[2];
    c[1][1] = mom[6] / n - mu[1] * mu[1];
    c[1][2] = c[2][1] = mom[7] / n - mu[1] * mu[2];
    c[2][2] = mom[8] / n - mu[2] * mu[2];
    double m = 0.0, v = 0.0;
    for (int a = 0; a < k; ++a) {
        double wa = (double)W0row[a];
        m += wa * mu[keepd[a]];
        for (int b = 0; b < k; ++b) v += wa * (double)W0row[b] * c[keepd[a]][keepd[b]];
    }
    mean0 = (float)m;
    var0 = (float)fmax(v, 0.0);
}

// Per-channel vectors (q0, r0, mi1, w2, b2, ...) of one layer for any struct with those members and
// FPV entries per net.
template <int FPV, bool BWD, class WT, class WBT>
__device__ __forceinline__ void stage_vectors(WT& W, WBT* WB, const float* raw, const LayerSrc& src, int F,
                                              unsigned wmask, bool train, bool phase0, const double* bsum1, int tid,
                                              int nthreads) {
    const int w = popc3(wmask), k = 3 - w;
    const NetOffsets o = net_offsets(F, w);
    int keepd[3], warpd[3];
    {
        int a = 0, b = 0;
        for (int d = 0; d < 3; ++d) {
            if (wmask & (1u << d)) warpd[b++] = d; else keepd[a++] = d;
        }
    }
    const int rs = rec_stride_of(F);
    const float* bn = raw + rs;
    for (int i = tid; i < 2 * FPV; i += nthreads) {
        const int net = i / FPV, c = i - net * FPV;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 r = q;
        float4 w2 = q;
        float2 mi = make_float2(0.f, 0.f), ab = mi;
        float g0 = 0.f;
        if (c < F) {
            const float* P = raw + net * o.stride;
            float mean0, var0;
            if (train) bn0_from_moments(src.mom, src.n_total, P + o.W0 + c * k, k, keepd, mean0, var0);
            else { mean0 = bn[net * 4 * F + c]; var0 = bn[net * 4 * F + F + c]; }
            const float i0 = 1.0f / sqrtf(var0 + GWTF_BN_EPS);
            g0 = P[o.g0 + c];
            const float b0 = P[o.b0 + c];
            float qa[3] = {0.f, 0.f, 0.f}, ra[3] = {0.f, 0.f, 0.f};
            for (int a = 0; a < k; ++a) {
                const float wv = P[o.W0 + c * k + a];
                ra[keepd[a]] = i0 * wv;
                qa[keepd[a]] = g0 * i0 * wv;
            }
            q = make_float4(qa[0], qa[1], qa[2], b0 - g0 * i0 * mean0);
            r = make_float4(ra[0], ra[1], ra[2], -i0 * mean0);
            if (!phase0) {
                float mean1, var1;
                if (train) {
                    const double s1 = src.sum1[(net * 2 + 0) * F + c], s2 = src.sum1[(net * 2 + 1) * F + c];
                    const double m = s1 / src.n_total;
                    mean1 = (float)m;
                    var1 = (float)fmax(s2 / src.n_total - m * m, 0.0);
                } else { mean1 = bn[net * 4 * F + 2 * F + c]; var1 = bn[net * 4 * F + 3 * F + c]; }
                const float i1 = 1.0f / sqrtf(var1 + GWTF_BN_EPS);
                float wa[3] = {0.f, 0.f, 0.f};
                for (int a = 0; a < w; ++a) wa[warpd[a]] = P[o.W2 + a * F + c];
                w2 = make_float4(wa[0], wa[1], wa[2], 0.f);
                mi = make_float2(mean1 * i1, i1);
                if (BWD && train && bsum1 != nullptr) {
                    ab = make_float2((float)(bsum1[(net * 4 + 0) * F + c] / src.n_total),
                                     (float)(bsum1[(net * 4 + 1) * F + c] / src.n_total));
                }
            }
        }
        W.q0[net][c] = q;
        W.w2[net][c] = w2;
        W.mi1[net][c] = mi;
        if (BWD) { WB->r0[net][c] = r; WB->sg[net][c].y = g0; WB->ab1[net][c] = ab; }
    }
    if (tid < 2) {
        const float* P = raw + tid * o.stride;
        float ba[3] = {0.f, 0.f, 0.f};
        if (!phase0) for (int a = 0; a < w; ++a) ba[warpd[a]] = P[o.b2 + a];
        W.b2[tid] = make_float4(ba[0], ba[1], ba[2], 0.f);
    }
}

// Build the compute layout from the raw record.  All threads of the CTA participate; caller
// syncs before (raw complete) and after (W ready).
//   train: BN statistics come from mom / sum1 (batch stats), else from the raw bn block.
//   phase0: only q0 + W1T are needed (statistics pass).
// sd1 weight, transposed + zero padded (depends on the raw record only)
template <int FP>
__device__ __forceinline__ void stage_w1t(LayerW<FP>& W, const float* raw, int F, unsigned wmask, int tid, int nthreads) {
    const NetOffsets o = net_offsets(F, popc3(wmask));
    for (int i = tid; i < 2 * FP * FP; i += nthreads) {
        const int net = i / (FP * FP), rem = i - net * FP * FP;
        const int f = rem / FP, e = rem - f * FP;   // read order: e fastest (coalesced in raw)
        float v = 0.f;
        if (f < F && e < F) v = raw[net * o.stride + o.W1 + f * F + e];
        W.W1T[net][e][f] = v;
    }
}
template <int FP, bool BWD>
__device__ __forceinline__ void stage_layer(LayerW<FP>& W, LayerWB<FP>* WB, const float* raw, const LayerSrc& src,
                                            int F, unsigned wmask, bool train, bool phase0, const double* bsum1,
                                            int tid, int nthreads) {
    stage_vectors<FP, BWD>(W, WB, raw, src, F, wmask, train, phase0, bsum1, tid, nthreads);
    stage_w1t<FP>(W, raw, F, wmask, tid, nthreads);
}

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization attribute
// may start while its predecessor in the stream is still draining; everything it does before pdl_wait()
// must depend on data older than the predecessor only (here: the parameter record and what is built from
// it).  pdl_trigger() lets the NEXT kernel start launching.  Both are no-ops in an ordinary launch.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// FiLM fold for one shape: st = (s*istd1, t - s*mean1*istd1).  `film` = 4F floats (smem or global):
// s_mu | t_mu | s_lv | t_lv.  Needs W.mi1 (stage_layer) to be visible.
template <int FP, bool BWD, class WT, class WBT>
__device__ __forceinline__ void stage_film(WT& W, WBT* WB, const float* film, int F, int tid, int nthreads) {
    for (int i = tid; i < 2 * FP; i += nthreads) {
        const int net = i / FP, c = i - net * FP;
        float2 st = make_float2(0.f, 0.f);
        float s = 0.f;
        if (c < F) {
            s = film[net * 2 * F + c];
            const float t = film[net * 2 * F + F + c];
            const float2 mi = W.mi1[net][c];
            st = make_float2(s * mi.y, t - s * mi.x);
        }
        W.st[net][c] = st;
        if (BWD) WB->sg[net][c].x = s;
    }
}

// The same fold with the 4F FiLM values of the NEXT shape already in registers (one (s, t) pair in each of the first 2 FP
// threads): a persistent kernel restages per shape behind a barrier of all its compute warps, and the global-load latency
// of these few values would otherwise be paid there, by everybody, at every shape boundary.
template <int FP>
struct FilmAhead {
    float s, t;
    int b;                  // shape the pair belongs to
    __device__ __forceinline__ void fetch(const float* film_all, int bb, int B, int K, int j, int L, int l, int F, int tid) {
        b = bb; s = 0.f; t = 0.f;
        if (tid < 2 * FP && bb < B) {
            const int net = tid / FP, c = tid - net * FP;
            if (c < F) {
                const float* f = film_all + ((size_t)(bb * K + j) * L + l) * 4 * F + net * 2 * F + c;
                s = f[0];
                t = f[F];
            }
        }
    }
    template <bool BWD, class WT, class WBT>
    __device__ __forceinline__ void stage(WT& W, WBT* WB, int F, int tid) const {
        if (tid < 2 * FP) {
            const int net = tid / FP, c = tid - net * FP;
            float2 st = make_float2(0.f, 0.f);
            if (c < F) {
                const float2 mi = W.mi1[net][c];
                st = make_float2(s * mi.y, t - s * mi.x);
            }
            W.st[net][c] = st;
            if (BWD) WB->sg[net][c].x = s;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// per-point microkernel
// ---------------------------------------------------------------------------------------------
// h1[p][f] = sum_e W1[f][e] * relu(q0[e] . (x,1)) for P points held by this thread.
template <int FP, int P>
__device__ __forceinline__ void contract_h1(const LayerW<FP>& W, int net, int F, const float (&x)[P][3],
                                            float (&acc)[P][FP]) {
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
        for (int f = 0; f < FP; ++f) acc[p][f] = 0.f;
    const float4* q0 = W.q0[net];
    const float4* w1 = reinterpret_cast<const float4*>(&W.W1T[net][0][0]);
#pragma unroll 1
    for (int e = 0; e < F; ++e) {
        const float4 q = q0[e];
        float a[P];
#pragma unroll
        for (int p = 0; p < P; ++p)
            a[p] = fmaxf(fmaf(q.x, x[p][0], fmaf(q.y, x[p][1], fmaf(q.z, x[p][2], q.w))), 0.f);
#pragma unroll
        for (int f4 = 0; f4 < FP / 4; ++f4) {
            const float4 wv = w1[e * (FP / 4) + f4];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                acc[p][4 * f4 + 0] = fmaf(wv.x, a[p], acc[p][4 * f4 + 0]);
                acc[p][4 * f4 + 1] = fmaf(wv.y, a[p], acc[p][4 * f4 + 1]);
                acc[p][4 * f4 + 2] = fmaf(wv.z, a[p], acc[p][4 * f4 + 2]);
                acc[p][4 * f4 + 3] = fmaf(wv.w, a[p], acc[p][4 * f4 + 3]);
            }
        }
    }
}

// o[p][d] = b2[d] + sum_f w2[f][d] * relu(st.x*h1 + st.y)
template <int FP, int P>
__device__ __forceinline__ void head_out(const LayerW<FP>& W, int net, const float (&acc)[P][FP], float (&o)[P][3]) {
    const float4 b = W.b2[net];
#pragma unroll
    for (int p = 0; p < P; ++p) { o[p][0] = b.x; o[p][1] = b.y; o[p][2] = b.z; }
#pragma unroll
    for (int f = 0; f < FP; ++f) {
        const float2 st = W.st[net][f];
        const float4 w2 = W.w2[net][f];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float a1 = fmaxf(fmaf(st.x, acc[p][f], st.y), 0.f);
            o[p][0] = fmaf(w2.x, a1, o[p][0]);
            o[p][1] = fmaf(w2.y, a1, o[p][1]);
            o[p][2] = fmaf(w2.z, a1, o[p][2]);
        }
    }
}

// y1 = st.x*h1 + st.y of both nets is kept from the apply pass ([net][f][b][n], coalesced over n) so the
// backward phases need not recompute the F x F contraction: h1 = (y1 - st.y) / st.x.
template <int FP, int P, int TT, class WT>
__device__ __forceinline__ void store_y1(const WT& W, int net, int F, const float (&acc)[P][FP], float* y1, int B,
                                         int N, int b, int n0, int tid, const bool (&valid)[P]) {
#pragma unroll
    for (int f = 0; f < FP; ++f) {
        if (f < F) {
            const float2 st = W.st[net][f];
            float* dst = y1 + (((size_t)net * F + f) * B + b) * N;
#pragma unroll
            for (int p = 0; p < P; ++p)
                if (valid[p]) dst[n0 + p * TT + tid] = fmaf(st.x, acc[p][f], st.y);
        }
    }
}
template <int FP, int P, int TT, class WT>
__device__ __forceinline__ void load_h1(const WT& W, int net, int F, float (&acc)[P][FP], const float* y1, int B, int N,
                                        int b, int n0, int tid, const bool (&valid)[P]) {
    // every load is unconditional on a clamped (always valid) address so that all P*FP of them are
    // in flight together; padding channels / points are zeroed afterwards
    const float* base = y1 + ((size_t)net * F * B + b) * N;
    const size_t fstride = (size_t)B * N;
    int idx[P];
#pragma unroll
    for (int p = 0; p < P; ++p) idx[p] = min(n0 + p * TT + tid, N - 1);
#pragma unroll
    for (int f = 0; f < FP; ++f) {
        const float* src = base + (size_t)min(f, F - 1) * fstride;
#pragma unroll
        for (int p = 0; p < P; ++p) acc[p][f] = __ldg(src + idx[p]);
    }
#pragma unroll
    for (int f = 0; f < FP; ++f) {
        const float2 st = W.st[net][f];
        const float inv = f < F ? 1.0f / st.x : 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p) acc[p][f] = (valid[p] && f < F) ? (acc[p][f] - st.y) * inv : 0.f;
    }
}

__device__ __forceinline__ float softsign(float v) { return v / (1.0f + fabsf(v)); }

// flows.py:113/115 on all three dims (kept dims see mu = logvar = 0 exactly).
template <bool DIRECT>
__device__ __forceinline__ void warp_point(float (&x)[3], const float (&omu)[3], const float (&olv)[3], float (&lam)[3]) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        lam[d] = softsign(olv[d]);
        const float sig = sqrtf(GWTF_FLOW_EPS + expf(lam[d]));
        x[d] = DIRECT ? fmaf(sig, x[d], omu[d]) : (x[d] - omu[d]) / sig;
    }
}

// ---------------------------------------------------------------------------------------------
// warp reduce-scatter of 32 per-lane values: lane L returns sum over lanes of v[L]  (31 shuffles)
// ---------------------------------------------------------------------------------------------
template <int HALF>
struct RSStep {
    __device__ static __forceinline__ void run(float (&v)[32], int lane) {
        const bool up = (lane & HALF) != 0;
#pragma unroll
        for (int i = 0; i < HALF; ++i) {
            const float send = up ? v[i] : v[i + HALF];
            const float keep = up ? v[i + HALF] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, HALF);
        }
        RSStep<HALF / 2>::run(v, lane);
    }
};
template <>
struct RSStep<0> {
    __device__ static __forceinline__ void run(float (&)[32], int) {}
};
__device__ __forceinline__ float warp_reduce_scatter32(float (&v)[32], int lane) {
    RSStep<16>::run(v, lane);
    return v[0];
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }
__device__ __forceinline__ float u01_open(uint32_t x) { return ((float)(x >> 8) + 0.5f) * 5.9604644775390625e-08f; }

}  // namespace gwtf
