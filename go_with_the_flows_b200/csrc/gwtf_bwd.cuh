// Backward kernels: seeds from the mixture NLL, two phases per coupling layer
// (activations are recomputed from the saved layer inputs, never stored), closed-form finish.
#pragma once
#include "gwtf_common.cuh"
#include "gwtf_exchange.cuh"
#include "gwtf_mma.cuh"

namespace gwtf {

// =============================================================================================
// Seeds: responsibilities r_j = softmax_j(logp_j + logw_j), dL/dz, dL/dS, base / weight grads.
// grid (tiles, B): a CTA only sees points of one shape so per-shape sums reduce in-block.
// =============================================================================================
static __global__ void __launch_bounds__(kThreads) k_bwd_seed(int K, int B, int N, const float* __restrict__ z,
                                                       const float* __restrict__ ld, const float* __restrict__ base,
                                                       const float* __restrict__ logw, const float* __restrict__ nll,
                                                       const float* __restrict__ dnll, float* gbuf, float* gs,
                                                       float* dbase, float* dlogw) {
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
    float mu[3], lv[3], iv[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { mu[d] = base[b * 6 + d]; lv[d] = base[b * 6 + 3 + d]; iv[d] = 1.0f / expf(lv[d]); }
    for (int n = blockIdx.x * kThreads + tid; n < N; n += gridDim.x * kThreads) {
        const float nl = nll[(size_t)b * N + n], dn = dnll[(size_t)b * N + n];
#pragma unroll 1
        for (int j = 0; j < K; ++j) {
            const size_t o3 = ((size_t)j * B + b) * 3 * N + n;
            float dz[3], tot = ld[((size_t)j * B + b) * N + n];
#pragma unroll
            for (int d = 0; d < 3; ++d) { dz[d] = z[o3 + (size_t)d * N] - mu[d]; tot += lv[d] + dz[d] * dz[d] * iv[d]; }
            const float lp = -0.5f * (tot + 3.0f * GWTF_LOG_2PI);
            const float r = expf(lp + logw[b * K + j] + nl) * dn;     // dnll * responsibility
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float gz = r * dz[d] * iv[d];
                gbuf[o3 + (size_t)d * N] = gz;
                gs[o3 + (size_t)d * N] = 0.5f * r;
                v[d] -= gz;                                           // d mu_base
                v[3 + d] += 0.5f * r * (1.0f - dz[d] * dz[d] * iv[d]); // d logvar_base
            }
            // d logw[j]: K <= 16 slots after the 6 base slots (compile-time indices via select)
#pragma unroll
            for (int jj = 0; jj < GWTF_MAX_COMPONENTS; ++jj) v[6 + jj] -= (jj == j) ? r : 0.f;
        }
    }
    const float s = warp_reduce_scatter32(v, lane);
    if (lane < 6) atomicAdd(&dbase[b * 6 + lane], s);
    else if (lane < 6 + K) atomicAdd(&dlogw[b * K + lane - 6], s);
}

// =============================================================================================
// per-layer backward
// =============================================================================================
struct BwdArgs {
    gwtf_stack_desc d;
    int layer, train;
    const float *params, *bnbuf, *film;
    const float* xin;        // input of the layer: (K,B,3,N) or the (B,3,N) data cloud
    int xin_shared;
    const float* xout;       // output of the layer = ubuf[layer] (K,B,3,N)
    const float* y1in;       // kept by the forward apply pass (layout private to the engine), or null: recompute h1
    int kept_y1;             // mma kernels: the kept buffer holds y1 = s*n1 + t (tcgen05 forward) instead of h1
    const double* mom_in;    // (K,16) moments of the layer input
    const double* sum1;      // (K,2,2,F)
    double* bsum;            // (K,2,4,F) this layer: sum dn1 | sum dn1*n1 | sum dy0 | sum dy0*hhat0
    // lazy bn0 correction of the layer processed before this one (layer-1): G -= M*x - c
    const double* mom_prev;  // (K,16) input moments of layer-1 (null: no correction)
    const double* bsum_prev; // (K,2,4,F) of layer-1
    float* gbuf;             // (K,B,3,N) in: dL/d(out) (uncorrected)  out: dL/d(in) partial
    const float* gs;         // (K,B,3,N) dL/dS
    float* dobuf;            // (K,B,6,N) d o_mu | d o_lv
    float* dparams;          // layout of params (+=)
    float* dfilm;            // layout of film (+=)
    int B, N, tiles_per_shape;
    double n_total;
    // multi-rank train mode, tcgen05 kernels: the exchange of the sums this launch completes, run by its last CTA
    ExchangeTail tail;
};

// (M, c) of the lazy bn0 correction for component j of layer `lc`:  G_in -= M x - c.
// Computed by the first 2*FP threads, reduced into corr[12] (shared, pre-zeroed).
template <int FP>
__device__ __forceinline__ void bn0_correction(const gwtf_stack_desc& d, const float* params, int j, int lc,
                                               const double* mom, const double* bsum, double n_total, float* corr,
                                               int tid) {
    const int F = d.n_features, L = d.n_layers;
    const unsigned wm = d.warp_mask[lc];
    const int w = popc3(wm), k = 3 - w;
    int keepd[3];
    { int q = 0; for (int dd = 0; dd < 3; ++dd) if (!(wm & (1u << dd))) keepd[q++] = dd; }
    const NetOffsets o = net_offsets(F, w);
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
    if (tid < 2 * FP) {
        const int net = tid / FP, e = tid - net * FP;
        if (e < F) {
            const float* P = params + (size_t)(j * L + lc) * d.rec_stride + net * o.stride;
            float mean0, var0;
            bn0_from_moments(mom, n_total, P + o.W0 + e * k, k, keepd, mean0, var0);
            const float i0 = 1.0f / sqrtf(var0 + GWTF_BN_EPS);
            const float g0 = P[o.g0 + e];
            const float A = g0 * (float)(bsum[(net * 4 + 2) * F + e] / n_total);
            const float Bm = g0 * (float)(bsum[(net * 4 + 3) * F + e] / n_total);
            float wf[3] = {0.f, 0.f, 0.f};
            for (int a = 0; a < k; ++a) wf[keepd[a]] = P[o.W0 + e * k + a];
            const float i2b = i0 * i0 * Bm;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 3; ++c) v[r * 3 + c] = wf[r] * wf[c] * i2b;
                v[9 + r] = wf[r] * (i2b * mean0 - i0 * A);
            }
        }
    }
    if (tid < round_up(2 * FP, 32)) {
        const float s = warp_reduce_scatter32(v, tid & 31);
        if ((tid & 31) < 12) atomicAdd(&corr[tid & 31], s);
    }
}

template <int FP>
struct BwdSmem {
    LayerW<FP> W;
    LayerWB<FP> WB;
    uint64_t bar;
    float corr[12];
    float red[2][round_up(5 * FP + 3, 32)];
    uint32_t tmem_base;     // mma phase 0: tensor-memory accumulators
    float2 mif[2][FP];      // mma phase 0: n1 = v*mif.y - mif.x for the value v read/recomputed (h1 or kept y1)
};

// ---- PHASE 0: d(o_mu,o_lv), FiLM + sd2 gradients, sd1_bn backward sums ------------------------
// h1 of the thread's P points is stashed in shared memory ([f][p][tid], conflict-free) right after
// the contraction so that the 5-sums-per-channel reduction is a ROLLED loop over groups of 6
// channels (30 values per warp reduce-scatter): small code, no spills, and P = 4 points per thread.
template <int FP, int P>
__global__ void __launch_bounds__(kThreads) k_bwd_layer_d(const BwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdSmem<FP>& S = *reinterpret_cast<BwdSmem<FP>*>(smem_raw);
    float* raw = reinterpret_cast<float*>(smem_raw + round_up((int)sizeof(BwdSmem<FP>), 16));
    const int F = a.d.n_features, K = a.d.n_components, L = a.d.n_layers;
    float* hbuf = raw + round_up(raw_floats(F), 4);             // [FP][P][kThreads]
    const int tid = threadIdx.x, lane = tid & 31;
    const int j = blockIdx.y, l = a.layer;
    const int N = a.N, B = a.B;
    const bool train = a.train != 0;
    constexpr int NV = 5 * FP + 3;               // per net: (ds, dt, dW2 x3) per f, db2 x3
    constexpr int FG = 6;                        // channels per reduce-scatter group
    constexpr int NGF = (FP + FG - 1) / FG;

    LayerSrc src;
    src.params = a.params + (size_t)(j * L + l) * a.d.rec_stride;
    src.bn = a.bnbuf + (size_t)(j * L + l) * 8 * F;
    src.film = nullptr;
    src.mom = a.mom_in ? a.mom_in + j * GWTF_MOM_STRIDE : nullptr;
    src.sum1 = a.sum1 ? a.sum1 + (size_t)j * 4 * F : nullptr;
    src.n_total = a.n_total;

    if (tid == 0) { mbar_init(&S.bar, 1); mbar_fence_init(); }
    if (tid < 12) S.corr[tid] = 0.f;
    for (int i = tid; i < 2 * round_up(NV, 32); i += kThreads) (&S.red[0][0])[i] = 0.f;
    __syncthreads();
    if (tid == 0) issue_layer_copy(raw, src, F, !train, false, &S.bar);
    const bool correct = train && a.mom_prev != nullptr;
    if (correct)
        bn0_correction<FP>(a.d, a.params, j, l - 1, a.mom_prev + j * GWTF_MOM_STRIDE, a.bsum_prev + (size_t)j * 8 * F,
                           a.n_total, S.corr, tid);
    mbar_wait(&S.bar, 0u);
    stage_layer<FP, true>(S.W, &S.WB, raw, src, F, a.d.warp_mask[l], train, false, nullptr, tid, kThreads);
    __syncthreads();
    float M[9], cc[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) M[i] = correct ? S.corr[i] : 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) cc[i] = correct ? S.corr[9 + i] : 0.f;

    // natural-layout targets of the sd2 gradients
    const unsigned wm = a.d.warp_mask[l];
    const int w = popc3(wm);
    const NetOffsets o = net_offsets(F, w);
    int row_of_dim[3];
    { int q = 0; for (int dd = 0; dd < 3; ++dd) row_of_dim[dd] = (wm & (1u << dd)) ? q++ : -1; }
    int cur_b = -1;

    auto flush = [&](int b) {
        // block partials -> FiLM grads of shape b, sd2 grads, sd1_bn backward sums
        __syncthreads();
        float* dfl = a.dfilm + ((size_t)(b * K + j) * L + l) * 4 * F;
        float* dpr = a.dparams + (size_t)(j * L + l) * a.d.rec_stride;
        double* bs = a.bsum + (size_t)j * 8 * F;
        for (int i = tid; i < 2 * NV; i += kThreads) {
            const int net = i / NV, idx = i - net * NV;
            const float val = S.red[net][idx];
            if (idx < 5 * FP) {
                const int f = idx / 5, c = idx - 5 * f;
                if (f < F) {
                    if (c == 0) {            // ds
                        atomicAdd(&dfl[net * 2 * F + f], val);
                        if (train) atomicAdd(&bs[(net * 4 + 1) * F + f], (double)(val * S.WB.sg[net][f].x));
                    } else if (c == 1) {     // dt
                        atomicAdd(&dfl[net * 2 * F + F + f], val);
                        if (train) atomicAdd(&bs[(net * 4 + 0) * F + f], (double)(val * S.WB.sg[net][f].x));
                    } else {
                        const int row = row_of_dim[c - 2];
                        if (row >= 0) atomicAdd(&dpr[net * o.stride + o.W2 + row * F + f], val);
                    }
                }
            } else {
                const int row = row_of_dim[idx - 5 * FP];
                if (row >= 0) atomicAdd(&dpr[net * o.stride + o.b2 + row], val);
            }
        }
        __syncthreads();
        for (int i = tid; i < 2 * round_up(NV, 32); i += kThreads) (&S.red[0][0])[i] = 0.f;
        __syncthreads();
    };

    const int total_tiles = B * a.tiles_per_shape;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int b = t / a.tiles_per_shape;
        const int n0 = (t - b * a.tiles_per_shape) * (kThreads * P);
        if (b != cur_b) {
            if (cur_b >= 0) flush(cur_b);
            else __syncthreads();
            stage_film<FP, true>(S.W, &S.WB, a.film + ((size_t)(b * K + j) * L + l) * 4 * F, F, tid, kThreads);
            __syncthreads();
            cur_b = b;
        }
        float x[P][3], dom[P][3], dov[P][3];
        bool valid[P];
        const float* xin = a.xin_shared ? a.xin + (size_t)b * 3 * N : a.xin + ((size_t)j * B + b) * 3 * N;
        const size_t sb = ((size_t)j * B + b) * 3 * N;
        float Gc[P][3], out[P][3];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int n = n0 + p * kThreads + tid;
            valid[p] = n < N;
            float G[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                x[p][d] = valid[p] ? xin[(size_t)d * N + n] : 0.f;
                out[p][d] = valid[p] ? a.xout[sb + (size_t)d * N + n] : 0.f;
                G[d] = valid[p] ? a.gbuf[sb + (size_t)d * N + n] : 0.f;
            }
#pragma unroll
            for (int d = 0; d < 3; ++d)
                Gc[p][d] = G[d] - (M[d * 3 + 0] * out[p][0] + M[d * 3 + 1] * out[p][1] + M[d * 3 + 2] * out[p][2]) + cc[d];
        }
        // logvar net first: its output fixes sigma, hence both d o_mu and d o_lv
#pragma unroll 1
        for (int net = 1; net >= 0; --net) {
            {
                float acc[P][FP];
                if (a.y1in) load_h1<FP, P, kThreads>(S.W, net, F, acc, a.y1in + (size_t)j * 2 * F * B * N, B, N, b, n0, tid, valid);
                else contract_h1<FP, P>(S.W, net, F, x, acc);
                if (net == 1) {
                    float olv[P][3];
                    head_out<FP, P>(S.W, 1, acc, olv);
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const int n = n0 + p * kThreads + tid;
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            const float lam = softsign(olv[p][d]);
                            const float ex = expf(lam);
                            const float sig2 = GWTF_FLOW_EPS + ex;
                            const float sig = sqrtf(sig2);
                            const float gsd = valid[p] ? a.gs[sb + (size_t)d * N + n] : 0.f;
                            const float gd = valid[p] ? Gc[p][d] : 0.f;
                            const float gin = gd / sig;
                            dom[p][d] = -gin;
                            const float dlam = gsd - gd * out[p][d] * ex / (2.0f * sig2);
                            const float den = 1.0f + fabsf(olv[p][d]);
                            dov[p][d] = dlam / (den * den);
                            if (valid[p]) {
                                a.gbuf[sb + (size_t)d * N + n] = gin;
                                a.dobuf[((size_t)j * B + b) * 6 * N + (size_t)d * N + n] = dom[p][d];
                                a.dobuf[((size_t)j * B + b) * 6 * N + (size_t)(3 + d) * N + n] = dov[p][d];
                            }
                        }
                    }
                }
#pragma unroll
                for (int f = 0; f < FP; ++f)
#pragma unroll
                    for (int p = 0; p < P; ++p) hbuf[(f * P + p) * kThreads + tid] = acc[p][f];
            }
            float dO[P][3];
#pragma unroll
            for (int p = 0; p < P; ++p)
#pragma unroll
                for (int d = 0; d < 3; ++d) dO[p][d] = net == 1 ? dov[p][d] : dom[p][d];
            // 5 sums per channel, 6 channels (30 values) per warp reduce-scatter; rolled
#pragma unroll 1
            for (int g = 0; g < NGF; ++g) {
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
                for (int ff = 0; ff < FG; ++ff) {
                    const int f = g * FG + ff;
                    if (f < FP) {
                        const float2 st = S.W.st[net][f];
                        const float2 mi = S.W.mi1[net][f];
                        const float4 w2 = S.W.w2[net][f];
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            const float h = hbuf[(f * P + p) * kThreads + tid];
                            const float y1 = fmaf(st.x, h, st.y);
                            const float da1 = w2.x * dO[p][0] + w2.y * dO[p][1] + w2.z * dO[p][2];
                            const float dy1 = y1 > 0.f ? da1 : 0.f;
                            const float a1 = fmaxf(y1, 0.f);
                            v[ff * 5 + 0] += dy1 * fmaf(h, mi.y, -mi.x);
                            v[ff * 5 + 1] += dy1;
                            v[ff * 5 + 2] += dO[p][0] * a1;
                            v[ff * 5 + 3] += dO[p][1] * a1;
                            v[ff * 5 + 4] += dO[p][2] * a1;
                        }
                    }
                }
                const float r = warp_reduce_scatter32(v, lane);
                if (lane < 5 * FG && g * 5 * FG + lane < 5 * FP) atomicAdd(&S.red[net][g * 5 * FG + lane], r);
            }
            {   // sd2 bias sums
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
                for (int p = 0; p < P; ++p) { v[0] += dO[p][0]; v[1] += dO[p][1]; v[2] += dO[p][2]; }
                const float r = warp_reduce_scatter32(v, lane);
                if (lane < 3) atomicAdd(&S.red[net][5 * FP + lane], r);
            }
        }
    }
    if (cur_b >= 0) flush(cur_b);
}

// warp-level tensor-core MMA (register fragments) for the weight-gradient GEMM of phase 1:
// D(16x8) += A(16x8, row) * B(8x8, col), tf32 inputs, fp32 accumulate.  fp32-grade accuracy from the
// 3xTF32 split of both operands.
__device__ __forceinline__ void mma_m16n8k8_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// ---- PHASE 1: sd1 / sd0 / bn0 gradients and the input gradient --------------------------------
template <int FP>
struct BwdESmem {
    LayerW<FP> W;
    LayerWB<FP> WB;
    uint64_t bar;
    float red[round_up(5 * FP, 32) + 32];
};

template <int FP, int P>
__global__ void __launch_bounds__(kThreads) k_bwd_layer_e(const BwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdESmem<FP>& S = *reinterpret_cast<BwdESmem<FP>*>(smem_raw);
    float* raw = reinterpret_cast<float*>(smem_raw + round_up((int)sizeof(BwdESmem<FP>), 16));
    const int F = a.d.n_features, K = a.d.n_components, L = a.d.n_layers;
    float* Dh = raw + round_up(raw_floats(F), 4);           // [P*kThreads][FP]
    float* A0s = Dh + (size_t)P * kThreads * FP;            // [P*kThreads][FP]
    const int tid = threadIdx.x, lane = tid & 31;
    const int j = blockIdx.y, l = a.layer;
    const int N = a.N, B = a.B;
    const bool train = a.train != 0;
    constexpr int EG = 4;                                   // e's per reduce-scatter group (5 sums each); 4 so
                                                            // that a0 is written as one conflict-free STS.128
    constexpr int NGE = (FP + EG - 1) / EG;
    // weight-gradient GEMM dW1[f][e] = sum_rows Dh[row][f] * A0s[row][e] on the tensor cores (mma.sync
    // m16n8k8 tf32, 3xTF32): each warp owns 1/8 of the tile's rows and the whole (MT*16) x (NT*8) output
    constexpr int MT = (FP + 15) / 16, NT = (FP + 7) / 8;
    constexpr int NWARPS = kThreads / 32;
    constexpr int ROWS = P * kThreads;
    static_assert(ROWS % (NWARPS * 8) == 0, "rows per warp must be a multiple of the MMA K");

    LayerSrc src;
    src.params = a.params + (size_t)(j * L + l) * a.d.rec_stride;
    src.bn = a.bnbuf + (size_t)(j * L + l) * 8 * F;
    src.film = nullptr;
    src.mom = a.mom_in ? a.mom_in + j * GWTF_MOM_STRIDE : nullptr;
    src.sum1 = a.sum1 ? a.sum1 + (size_t)j * 4 * F : nullptr;
    src.n_total = a.n_total;

    if (tid == 0) { mbar_init(&S.bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (tid == 0) issue_layer_copy(raw, src, F, !train, false, &S.bar);
    mbar_wait(&S.bar, 0u);
    stage_layer<FP, true>(S.W, &S.WB, raw, src, F, a.d.warp_mask[l], train, false,
                          train ? a.bsum + (size_t)j * 8 * F : nullptr, tid, kThreads);
    __syncthreads();

    const unsigned wm = a.d.warp_mask[l];
    const int w = popc3(wm), k = 3 - w;
    const NetOffsets o = net_offsets(F, w);
    int col_of_dim[3];
    { int q = 0; for (int dd = 0; dd < 3; ++dd) col_of_dim[dd] = (wm & (1u << dd)) ? -1 : q++; }
    float* dpr = a.dparams + (size_t)(j * L + l) * a.d.rec_stride;
    double* bs = a.bsum + (size_t)j * 8 * F;

    const int warp = tid >> 5, fg = lane >> 2, ft = lane & 3;      // MMA fragment coordinates
    const int total_tiles = B * a.tiles_per_shape;

#pragma unroll 1
    for (int net = 0; net < 2; ++net) {
        float gacc[MT][NT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) gacc[mt][nt][i] = 0.f;
        float eacc[NGE];
#pragma unroll
        for (int g = 0; g < NGE; ++g) eacc[g] = 0.f;
        int cur_b = -1;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int b = t / a.tiles_per_shape;
            const int n0 = (t - b * a.tiles_per_shape) * (kThreads * P);
            if (b != cur_b) {
                __syncthreads();
                stage_film<FP, true>(S.W, &S.WB, a.film + ((size_t)(b * K + j) * L + l) * 4 * F, F, tid, kThreads);
                __syncthreads();
                cur_b = b;
            }
            float x[P][3], dO[P][3];
            bool valid[P];
            const float* xin = a.xin_shared ? a.xin + (size_t)b * 3 * N : a.xin + ((size_t)j * B + b) * 3 * N;
            const size_t sb = ((size_t)j * B + b) * 3 * N;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const int n = n0 + p * kThreads + tid;
                valid[p] = n < N;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    x[p][d] = valid[p] ? xin[(size_t)d * N + n] : 0.f;
                    dO[p][d] = valid[p] ? a.dobuf[((size_t)j * B + b) * 6 * N + (size_t)(net * 3 + d) * N + n] : 0.f;
                }
            }
            float acc[P][FP];
            if (a.y1in) load_h1<FP, P, kThreads>(S.W, net, F, acc, a.y1in + (size_t)j * 2 * F * B * N, B, N, b, n0, tid, valid);
            else contract_h1<FP, P>(S.W, net, F, x, acc);
            // h1 -> dh1 in place
#pragma unroll
            for (int f = 0; f < FP; ++f) {
                const float2 st = S.W.st[net][f];
                const float2 mi = S.W.mi1[net][f];
                const float4 w2 = S.W.w2[net][f];
                const float s = S.WB.sg[net][f].x;
                const float2 ab = S.WB.ab1[net][f];
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const float h = acc[p][f];
                    const float y1 = fmaf(st.x, h, st.y);
                    const float da1 = w2.x * dO[p][0] + w2.y * dO[p][1] + w2.z * dO[p][2];
                    const float dn1 = y1 > 0.f ? da1 * s : 0.f;
                    const float n1 = fmaf(h, mi.y, -mi.x);
                    acc[p][f] = valid[p] ? mi.y * (dn1 - ab.x - n1 * ab.y) : 0.f;
                }
            }
#pragma unroll
            for (int p = 0; p < P; ++p) {
                float4* row = reinterpret_cast<float4*>(Dh + (size_t)(p * kThreads + tid) * FP);
#pragma unroll
                for (int f4 = 0; f4 < FP / 4; ++f4)
                    row[f4] = make_float4(acc[p][4 * f4], acc[p][4 * f4 + 1], acc[p][4 * f4 + 2], acc[p][4 * f4 + 3]);
            }
            // da0 = W1^T dh1, bn0/sd0 pieces, input-gradient contribution
            float vin[P][3];
#pragma unroll
            for (int p = 0; p < P; ++p) vin[p][0] = vin[p][1] = vin[p][2] = 0.f;
            const float4* w1 = reinterpret_cast<const float4*>(&S.W.W1T[net][0][0]);
#pragma unroll 1
            for (int g = 0; g < NGE; ++g) {
                float v[32];
                float a0v[P][EG];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
                for (int ee = 0; ee < EG; ++ee) {
#pragma unroll
                    for (int p = 0; p < P; ++p) a0v[p][ee] = 0.f;
                    const int e = g * EG + ee;
                    if (e < FP) {
                        float da0[P];
#pragma unroll
                        for (int p = 0; p < P; ++p) da0[p] = 0.f;
#pragma unroll
                        for (int f4 = 0; f4 < FP / 4; ++f4) {
                            const float4 wv = w1[e * (FP / 4) + f4];
#pragma unroll
                            for (int p = 0; p < P; ++p) {
                                da0[p] = fmaf(wv.x, acc[p][4 * f4 + 0], da0[p]);
                                da0[p] = fmaf(wv.y, acc[p][4 * f4 + 1], da0[p]);
                                da0[p] = fmaf(wv.z, acc[p][4 * f4 + 2], da0[p]);
                                da0[p] = fmaf(wv.w, acc[p][4 * f4 + 3], da0[p]);
                            }
                        }
                        const float4 qv = S.W.q0[net][e];
                        const float4 rv = S.WB.r0[net][e];
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            const float y0 = fmaf(qv.x, x[p][0], fmaf(qv.y, x[p][1], fmaf(qv.z, x[p][2], qv.w)));
                            const float dy0 = y0 > 0.f ? da0[p] : 0.f;
                            const float hh = fmaf(rv.x, x[p][0], fmaf(rv.y, x[p][1], fmaf(rv.z, x[p][2], rv.w)));
                            a0v[p][ee] = valid[p] ? fmaxf(y0, 0.f) : 0.f;
                            vin[p][0] = fmaf(qv.x, dy0, vin[p][0]);
                            vin[p][1] = fmaf(qv.y, dy0, vin[p][1]);
                            vin[p][2] = fmaf(qv.z, dy0, vin[p][2]);
                            v[ee * 5 + 0] += dy0 * hh;       // d gamma0
                            v[ee * 5 + 1] += dy0;            // d beta0
                            v[ee * 5 + 2] += dy0 * x[p][0];  // raw sd0 weight sums
                            v[ee * 5 + 3] += dy0 * x[p][1];
                            v[ee * 5 + 4] += dy0 * x[p][2];
                        }
                    }
                }
                if (g * EG < FP) {
#pragma unroll
                    for (int p = 0; p < P; ++p)
                        *reinterpret_cast<float4*>(A0s + (size_t)(p * kThreads + tid) * FP + g * EG) =
                            make_float4(a0v[p][0], a0v[p][1], a0v[p][2], a0v[p][3]);
                }
                eacc[g] += warp_reduce_scatter32(v, lane);
            }
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const int n = n0 + p * kThreads + tid;
                if (valid[p])
#pragma unroll
                    for (int d = 0; d < 3; ++d) a.gbuf[sb + (size_t)d * N + n] += vin[p][d];
            }
            __syncthreads();
            // dW1[f][e] += sum_rows Dh[row][f] * A0s[row][e]   (A = Dh^T: m = f, k = row; B: k = row, n = e)
            {
                const int r0 = warp * (ROWS / NWARPS);
#pragma unroll 1
                for (int kk = 0; kk < ROWS / NWARPS; kk += 8) {
                    const float* dr0 = Dh + (size_t)(r0 + kk + ft) * FP;
                    const float* dr1 = dr0 + 4 * FP;
                    const float* ar0 = A0s + (size_t)(r0 + kk + ft) * FP;
                    const float* ar1 = ar0 + 4 * FP;
                    uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        split_tf32_bits(ar0[nt * 8 + fg], bh[nt][0], bl[nt][0]);
                        split_tf32_bits(ar1[nt * 8 + fg], bh[nt][1], bl[nt][1]);
                    }
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        uint32_t ah[4], al[4];
                        split_tf32_bits(dr0[mt * 16 + fg], ah[0], al[0]);
                        split_tf32_bits(dr0[mt * 16 + fg + 8], ah[1], al[1]);
                        split_tf32_bits(dr1[mt * 16 + fg], ah[2], al[2]);
                        split_tf32_bits(dr1[mt * 16 + fg + 8], ah[3], al[3]);
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) {
                            mma_m16n8k8_tf32(gacc[mt][nt], ah, bh[nt]);
                            mma_m16n8k8_tf32(gacc[mt][nt], al, bh[nt]);
                            mma_m16n8k8_tf32(gacc[mt][nt], ah, bl[nt]);
                        }
                    }
                }
            }
            __syncthreads();
        }
        // ---- flush the per-e sums of this net
        for (int i = tid; i < round_up(5 * FP, 32) + 32; i += kThreads) S.red[i] = 0.f;
        __syncthreads();
#pragma unroll
        for (int g = 0; g < NGE; ++g)
            if (lane < 5 * EG) atomicAdd(&S.red[g * 5 * EG + lane], eacc[g]);
        __syncthreads();
        for (int i = tid; i < 5 * FP; i += kThreads) {
            const int e = i / 5, c = i - 5 * e;
            if (e >= F) continue;
            const float val = S.red[i];
            if (c == 0) {
                atomicAdd(&dpr[net * o.stride + o.g0 + e], val);
                atomicAdd(&bs[(net * 4 + 3) * F + e], (double)val);
            } else if (c == 1) {
                atomicAdd(&dpr[net * o.stride + o.b0 + e], val);
                atomicAdd(&bs[(net * 4 + 2) * F + e], (double)val);
            } else {
                const int col = col_of_dim[c - 2];
                if (col >= 0) atomicAdd(&dpr[net * o.stride + o.W0 + e * k + col], val);
            }
        }
        // ---- flush dW1: reduce the per-warp partial outputs through shared memory (reuses Dh/A0s),
        // four warps at a time so the staging buffer stays within the operand area for every width
        constexpr int PM = MT * 16, PN = NT * 8;
        float* part = Dh;                                   // [4][MT*16][NT*8]
#pragma unroll 1
        for (int half = 0; half < NWARPS / 4; ++half) {
            __syncthreads();
            if ((warp >> 2) == half) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int f = mt * 16 + fg + ((i & 2) ? 8 : 0), e = nt * 8 + 2 * ft + (i & 1);
                            part[((size_t)(warp & 3) * PM + f) * PN + e] = gacc[mt][nt][i];
                        }
            }
            __syncthreads();
            for (int i = tid; i < F * F; i += kThreads) {
                const int f = i / F, e = i - f * F;
                float s = 0.f;
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) s += part[((size_t)qq * PM + f) * PN + e];
                atomicAdd(&dpr[net * o.stride + o.W1 + f * F + e], s);
            }
        }
        __syncthreads();
    }
}

// =============================================================================================
// Finish: closed-form sd0.weight gradients (bn0 backward folded in) and the data-point gradient
// =============================================================================================
struct FinishArgs {
    gwtf_stack_desc d;
    int train;
    const float *params, *bnbuf;
    const double *mom, *bsum;     // (L,K,16), (L,K,2,4,F)
    float* dparams;
    double n_total, local_frac;   // local_frac = this rank's share of the batch statistics
};

// dW0[e][a] = i0*g0*( P_raw - (dbeta0/n) Sx_a - (dgamma0/n) i0 (W0[e] . Sxx[:,a] - m0 Sx_a) )
static __global__ void k_bwd_finish_w0(const FinishArgs a) {
    const int F = a.d.n_features, K = a.d.n_components, L = a.d.n_layers;
    const int total = L * K * 2 * F;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int e = i % F, net = (i / F) % 2, j = (i / (2 * F)) % K, l = i / (2 * F * K);
        const unsigned wm = a.d.warp_mask[l];
        const int w = popc3(wm), k = 3 - w;
        int keepd[3];
        { int q = 0; for (int dd = 0; dd < 3; ++dd) if (!(wm & (1u << dd))) keepd[q++] = dd; }
        const NetOffsets o = net_offsets(F, w);
        const size_t rec = (size_t)(j * L + l) * a.d.rec_stride + net * o.stride;
        const float* P = a.params + rec;
        float* dP = a.dparams + rec;
        const float g0 = P[o.g0 + e];
        if (a.train) {
            const double* mom = a.mom + ((size_t)l * K + j) * GWTF_MOM_STRIDE;
            const double* bs = a.bsum + ((size_t)l * K + j) * 8 * F;
            float mean0, var0;
            bn0_from_moments(mom, a.n_total, P + o.W0 + e * k, k, keepd, mean0, var0);
            const double i0 = 1.0 / sqrt((double)var0 + (double)GWTF_BN_EPS);
            const double db = bs[(net * 4 + 2) * F + e] / a.n_total, dg = bs[(net * 4 + 3) * F + e] / a.n_total;
            // symmetric second-moment lookup
            auto sxx = [&](int r, int c) {
                if (r > c) { int t = r; r = c; c = t; }
                const int idx = r == 0 ? 3 + c : (r == 1 ? 5 + c : 8);
                return mom[idx];
            };
            for (int c = 0; c < k; ++c) {
                const int dc = keepd[c];
                double wsxx = 0.0;
                for (int r = 0; r < k; ++r) wsxx += (double)P[o.W0 + e * k + r] * sxx(keepd[r], dc);
                const double corr = db * mom[dc] + dg * i0 * (wsxx - (double)mean0 * mom[dc]);
                dP[o.W0 + e * k + c] = (float)(i0 * (double)g0 * ((double)dP[o.W0 + e * k + c] - a.local_frac * corr));
            }
        } else {
            const float* bn = a.bnbuf + (size_t)(j * L + l) * 8 * F + net * 4 * F;
            const float i0 = 1.0f / sqrtf(bn[F + e] + GWTF_BN_EPS);
            for (int c = 0; c < k; ++c) dP[o.W0 + e * k + c] *= i0 * g0;
        }
    }
}

// dpoints[b][d][n] = sum_j ( G_j - M_j x + c_j )   with the lazy bn0 correction of the last layer
template <int FP>
__global__ void __launch_bounds__(kThreads) k_bwd_finish_points(gwtf_stack_desc d, int train, const float* params,
                                                                const double* mom_last, const double* bsum_last,
                                                                const float* gbuf, const float* points,
                                                                float* dpoints, int B, int N, double n_total) {
    __shared__ float corr[GWTF_MAX_COMPONENTS][12];
    const int K = d.n_components, tid = threadIdx.x;
    for (int i = tid; i < GWTF_MAX_COMPONENTS * 12; i += kThreads) (&corr[0][0])[i] = 0.f;
    __syncthreads();
    if (train)
        for (int j = 0; j < K; ++j)
            bn0_correction<FP>(d, params, j, d.n_layers - 1, mom_last + j * GWTF_MOM_STRIDE,
                               bsum_last + (size_t)j * 8 * d.n_features, n_total, corr[j], tid);
    __syncthreads();
    const size_t total = (size_t)B * N;
    for (size_t i = (size_t)blockIdx.x * kThreads + tid; i < total; i += (size_t)gridDim.x * kThreads) {
        const int b = (int)(i / N), n = (int)(i - (size_t)b * N);
        float x[3], acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int dd = 0; dd < 3; ++dd) x[dd] = points[((size_t)b * 3 + dd) * N + n];
        for (int j = 0; j < K; ++j) {
            const float* M = corr[j];
#pragma unroll
            for (int dd = 0; dd < 3; ++dd) {
                const float g = gbuf[(((size_t)j * B + b) * 3 + dd) * N + n];
                acc[dd] += g - (M[dd * 3] * x[0] + M[dd * 3 + 1] * x[1] + M[dd * 3 + 2] * x[2]) + M[9 + dd];
            }
        }
#pragma unroll
        for (int dd = 0; dd < 3; ++dd) dpoints[((size_t)b * 3 + dd) * N + n] += acc[dd];
    }
}

}  // namespace gwtf
