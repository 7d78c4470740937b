// Forward kernels: fused eval-mode NLL, phased (batch-statistics) forward, NLL from state.
#pragma once
#include "gwtf_common.cuh"
#include "gwtf_exchange.cuh"

namespace gwtf {

// =============================================================================================
// Fused eval-mode forward NLL: one launch, every point walks K stacks x L layers in registers.
// replaces flow_mixture.py:163-166 (K x one_flow_decode, inverse) + losses.py:88-137.
// =============================================================================================
struct EvalArgs {
    gwtf_stack_desc d;
    const float *params, *bnbuf, *film, *points, *base, *logw;
    int B, N, tiles_per_shape;
    float *nll, *logp, *z, *ssum;
};

template <int FP>
struct EvalSmem {
    LayerW<FP> W;
    uint64_t bar[2];
};

template <int FP, int P>
__global__ void __launch_bounds__(kThreads) k_nll_eval(const EvalArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EvalSmem<FP>& S = *reinterpret_cast<EvalSmem<FP>*>(smem_raw);
    const int F = a.d.n_features, K = a.d.n_components, L = a.d.n_layers;
    const int RAW = round_up(raw_floats(F), 4);
    float* raw0 = reinterpret_cast<float*>(smem_raw + round_up((int)sizeof(EvalSmem<FP>), 16));
    float* rawb[2] = {raw0, raw0 + RAW};
    const int tid = threadIdx.x;
    const int b = blockIdx.x / a.tiles_per_shape;
    const int n0 = (blockIdx.x - b * a.tiles_per_shape) * (kThreads * P);
    const int N = a.N;

    if (tid == 0) { mbar_init(&S.bar[0], 1); mbar_init(&S.bar[1], 1); mbar_fence_init(); }
    __syncthreads();

    auto src_of = [&](int s) {
        const int j = s / L, l = L - 1 - (s - j * L);
        LayerSrc src;
        src.params = a.params + (size_t)(j * L + l) * a.d.rec_stride;
        src.bn = a.bnbuf + (size_t)(j * L + l) * 8 * F;
        src.film = a.film + ((size_t)(b * K + j) * L + l) * 4 * F;
        src.mom = nullptr; src.sum1 = nullptr; src.n_total = 1.0;
        return src;
    };
    const int total = K * L;
    if (tid == 0) issue_layer_copy(rawb[0], src_of(0), F, true, true, &S.bar[0]);

    float x0[P][3];
    bool valid[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int n = n0 + p * kThreads + tid;
        valid[p] = n < N;
#pragma unroll
        for (int d = 0; d < 3; ++d) x0[p][d] = valid[p] ? a.points[((size_t)b * 3 + d) * N + n] : 0.f;
    }
    float mub[3], lvb[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { mub[d] = a.base[b * 6 + d]; lvb[d] = a.base[b * 6 + 3 + d]; }

    float x[P][3], S3[P][3], lse_m[P], lse_s[P];
#pragma unroll
    for (int p = 0; p < P; ++p) { lse_m[p] = -INFINITY; lse_s[p] = 0.f; }

    uint32_t ph[2] = {0u, 0u};
    for (int s = 0; s < total; ++s) {
        const int j = s / L, l = L - 1 - (s - j * L);
        const int buf = s & 1;
        mbar_wait(&S.bar[buf], ph[buf]);
        ph[buf] ^= 1u;
        stage_layer<FP, false>(S.W, nullptr, rawb[buf], LayerSrc(), F, a.d.warp_mask[l], false, false, nullptr, tid,
                               kThreads);
        __syncthreads();
        stage_film<FP, false>(S.W, (LayerWB<FP>*)nullptr, rawb[buf] + rec_stride_of(F) + 8 * F, F, tid, kThreads);
        __syncthreads();
        if (tid == 0 && s + 1 < total) issue_layer_copy(rawb[buf ^ 1], src_of(s + 1), F, true, true, &S.bar[buf ^ 1]);

        if (l == L - 1) {
#pragma unroll
            for (int p = 0; p < P; ++p)
#pragma unroll
                for (int d = 0; d < 3; ++d) { x[p][d] = x0[p][d]; S3[p][d] = 0.f; }
        }
        {
            float omu[P][3], olv[P][3];
            {
                float acc[P][FP];
                contract_h1<FP, P>(S.W, 0, F, x, acc);
                head_out<FP, P>(S.W, 0, acc, omu);
            }
            {
                float acc[P][FP];
                contract_h1<FP, P>(S.W, 1, F, x, acc);
                head_out<FP, P>(S.W, 1, acc, olv);
            }
#pragma unroll
            for (int p = 0; p < P; ++p) {
                float lam[3];
                warp_point<false>(x[p], omu[p], olv[p], lam);
#pragma unroll
                for (int d = 0; d < 3; ++d) S3[p][d] += lam[d];
            }
        }
        if (l == 0) {
            const float lw = a.logw[b * K + j];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                float tot = 0.f;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float dz = x[p][d] - mub[d];
                    tot += (lvb[d] + S3[p][d]) + dz * dz / expf(lvb[d]);
                }
                const float lp = -0.5f * (tot + 3.0f * GWTF_LOG_2PI);
                const float v = lp + lw;
                if (v > lse_m[p]) { lse_s[p] = lse_s[p] * expf(lse_m[p] - v) + 1.0f; lse_m[p] = v; }
                else lse_s[p] += expf(v - lse_m[p]);
                const int n = n0 + p * kThreads + tid;
                if (valid[p]) {
                    if (a.logp) a.logp[((size_t)b * N + n) * K + j] = lp;
                    if (a.z)
#pragma unroll
                        for (int d = 0; d < 3; ++d) a.z[(((size_t)j * a.B + b) * 3 + d) * N + n] = x[p][d];
                    if (a.ssum)
#pragma unroll
                        for (int d = 0; d < 3; ++d) a.ssum[(((size_t)j * a.B + b) * 3 + d) * N + n] = S3[p][d];
                }
            }
        }
        __syncthreads();   // W is rebuilt next iteration
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int n = n0 + p * kThreads + tid;
        if (valid[p]) {
            const float out = -(lse_m[p] + logf(lse_s[p]));
            a.nll[(size_t)b * N + n] = out;
            if (a.d.nonfinite && !isfinite(out)) atomicAdd(a.d.nonfinite, 1);
        }
    }
}

// =============================================================================================
// Phased forward of ONE layer over all components (grid.y = K).
//   PHASE 0: statistics of h1 = W1 relu(bn0(W0 x))      -> sum1 (fp64 atomics)
//   PHASE 1: apply (flows.py:95-117), write output coords, add logvar, next-layer input moments
// =============================================================================================
struct LayerArgs {
    gwtf_stack_desc d;
    int layer, train, direct;
    const float *params, *bnbuf, *film;
    const float* xin;       // (K,B,3,N) or (B,3,N) when xin_shared
    int xin_shared;
    float* xout;            // (K,B,3,N)
    float* ld;              // (K,B,N) running sum of logvar over dims and layers (+=), may be null
    float* ssum;            // (K,B,3,N) per-dim running sum (+=), may be null
    float* trio;            // (K,3,B,3,N): p_out | mu | logvar of this layer, may be null
    float* y1out;           // (K,2,F,B,N): post-FiLM pre-activation of sd1 kept for backward, may be null
    const double* mom_in;   // (K,16) moments of this layer's input
    double* mom_out;        // (K,16) moments of this layer's output (+=), may be null
    double* sum1;           // (K,2,2,F)
    int B, N, tiles_per_shape;
    double n_total;
    // segmented mode (sampling through the tcgen05 layer kernels): xin / xout are (B,3,N) rows in which the points
    // that drew component j occupy [seg[j][b][0], + seg[j][b][1]) (segment starts are multiples of 128);
    // seg_tiles[j][0..B] = exclusive prefix of the segments' 128-point tile counts.  Null = dense (K,B,3,N).
    const int32_t* seg;
    const int32_t* seg_tiles;
    // multi-rank train mode, tcgen05 kernels: the exchange of the sums this launch completes, run by its last CTA
    ExchangeTail tail;
};

template <int FP>
struct PhaseSmem {
    LayerW<FP> W;
    uint64_t bar;
    float red[2][2 * FP + 32];
    double dred[16];
};

template <int FP, int P, int PHASE>
__global__ void __launch_bounds__(kThreads) k_fwd_layer(const LayerArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PhaseSmem<FP>& S = *reinterpret_cast<PhaseSmem<FP>*>(smem_raw);
    float* raw = reinterpret_cast<float*>(smem_raw + round_up((int)sizeof(PhaseSmem<FP>), 16));
    const int F = a.d.n_features, K = a.d.n_components, L = a.d.n_layers;
    const int tid = threadIdx.x, lane = tid & 31;
    const int j = blockIdx.y, l = a.layer;
    const int N = a.N, B = a.B;
    const bool train = a.train != 0;

    LayerSrc src;
    src.params = a.params + (size_t)(j * L + l) * a.d.rec_stride;
    src.bn = a.bnbuf + (size_t)(j * L + l) * 8 * F;
    src.film = nullptr;
    src.mom = a.mom_in ? a.mom_in + j * GWTF_MOM_STRIDE : nullptr;
    src.sum1 = a.sum1 ? a.sum1 + (size_t)j * 4 * F : nullptr;
    src.n_total = a.n_total;

    if (tid == 0) { mbar_init(&S.bar, 1); mbar_fence_init(); }
    for (int i = tid; i < 2 * (2 * FP + 32); i += kThreads) (&S.red[0][0])[i] = 0.f;
    if (tid < 16) S.dred[tid] = 0.0;
    __syncthreads();
    if (tid == 0) issue_layer_copy(raw, src, F, !train, false, &S.bar);
    mbar_wait(&S.bar, 0u);
    stage_layer<FP, false>(S.W, nullptr, raw, src, F, a.d.warp_mask[l], train, PHASE == 0, nullptr, tid, kThreads);
    __syncthreads();

    constexpr int NG = (2 * FP + 31) / 32;   // reduce-scatter groups for (h, h^2) per net
    float sacc[2][NG];
#pragma unroll
    for (int net = 0; net < 2; ++net)
#pragma unroll
        for (int g = 0; g < NG; ++g) sacc[net][g] = 0.f;
    float macc = 0.f;
    int cur_b = -1;

    const int total_tiles = B * a.tiles_per_shape;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int b = t / a.tiles_per_shape;
        const int n0 = (t - b * a.tiles_per_shape) * (kThreads * P);
        if (PHASE == 1 && b != cur_b) {
            __syncthreads();
            stage_film<FP, false>(S.W, (LayerWB<FP>*)nullptr, a.film + ((size_t)(b * K + j) * L + l) * 4 * F, F, tid, kThreads);
            __syncthreads();
            cur_b = b;
        }
        float x[P][3];
        bool valid[P];
        const float* xin = a.xin_shared ? a.xin + (size_t)b * 3 * N : a.xin + ((size_t)j * B + b) * 3 * N;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int n = n0 + p * kThreads + tid;
            valid[p] = n < N;
#pragma unroll
            for (int d = 0; d < 3; ++d) x[p][d] = valid[p] ? xin[(size_t)d * N + n] : 0.f;
        }
        if (PHASE == 0) {
#pragma unroll
            for (int net = 0; net < 2; ++net) {
                float acc[P][FP];
                contract_h1<FP, P>(S.W, net, F, x, acc);
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int idx = g * 32 + i, f = idx >> 1;
                        float s = 0.f;
                        if (f < FP) {
#pragma unroll
                            for (int p = 0; p < P; ++p) {
                                const float h = valid[p] ? acc[p][f] : 0.f;
                                s += (idx & 1) ? h * h : h;
                            }
                        }
                        v[i] = s;
                    }
                    sacc[net][g] += warp_reduce_scatter32(v, lane);
                }
            }
        } else {
            float omu[P][3], olv[P][3];
            {
                float acc[P][FP];
                contract_h1<FP, P>(S.W, 0, F, x, acc);
                head_out<FP, P>(S.W, 0, acc, omu);
                if (a.y1out) store_y1<FP, P, kThreads>(S.W, 0, F, acc, a.y1out + (size_t)j * 2 * F * B * N, B, N, b, n0, tid, valid);
            }
            {
                float acc[P][FP];
                contract_h1<FP, P>(S.W, 1, F, x, acc);
                head_out<FP, P>(S.W, 1, acc, olv);
                if (a.y1out) store_y1<FP, P, kThreads>(S.W, 1, F, acc, a.y1out + (size_t)j * 2 * F * B * N, B, N, b, n0, tid, valid);
            }
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                float lam[3];
                if (a.direct) warp_point<true>(x[p], omu[p], olv[p], lam);
                else warp_point<false>(x[p], omu[p], olv[p], lam);
                const int n = n0 + p * kThreads + tid;
                if (valid[p]) {
                    const size_t base = ((size_t)j * B + b) * 3 * N + n;
#pragma unroll
                    for (int d = 0; d < 3; ++d) a.xout[base + (size_t)d * N] = x[p][d];
                    if (a.ld) a.ld[((size_t)j * B + b) * N + n] += lam[0] + lam[1] + lam[2];
                    if (a.ssum)
#pragma unroll
                        for (int d = 0; d < 3; ++d) a.ssum[base + (size_t)d * N] += lam[d];
                    if (a.trio) {
                        const size_t tb = (((size_t)j * 3) * B + b) * 3 * N + n;
                        const size_t ts = (size_t)B * 3 * N;
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            a.trio[tb + (size_t)d * N] = x[p][d];
                            a.trio[tb + ts + (size_t)d * N] = omu[p][d];
                            a.trio[tb + 2 * ts + (size_t)d * N] = lam[d];
                        }
                    }
                    v[0] += x[p][0]; v[1] += x[p][1]; v[2] += x[p][2];
                    v[3] += x[p][0] * x[p][0]; v[4] += x[p][0] * x[p][1]; v[5] += x[p][0] * x[p][2];
                    v[6] += x[p][1] * x[p][1]; v[7] += x[p][1] * x[p][2]; v[8] += x[p][2] * x[p][2];
                }
            }
            if (a.mom_out) macc += warp_reduce_scatter32(v, lane);
        }
    }
    // ---- flush block partials
    if (PHASE == 0) {
#pragma unroll
        for (int net = 0; net < 2; ++net)
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const int idx = g * 32 + lane;
                if (idx < 2 * FP) atomicAdd(&S.red[net][idx], sacc[net][g]);
            }
        __syncthreads();
        for (int i = tid; i < 2 * 2 * FP; i += kThreads) {
            const int net = i / (2 * FP), idx = i - net * 2 * FP, f = idx >> 1, which = idx & 1;
            if (f < F) atomicAdd(&a.sum1[((size_t)j * 2 + net) * 2 * F + which * F + f], (double)S.red[net][idx]);
        }
    } else if (a.mom_out) {
        if (lane < 9) atomicAdd(&S.dred[lane], (double)macc);
        __syncthreads();
        if (tid < 9) atomicAdd(&a.mom_out[j * GWTF_MOM_STRIDE + tid], S.dred[tid]);
    }
}

// moments of the raw data points (input of the first processed layer), replicated to K slots
static __global__ void __launch_bounds__(kThreads) k_moments(const float* __restrict__ pts, int B, int N, int K, double* mom) {
    __shared__ double dred[16];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 16) dred[tid] = 0.0;
    __syncthreads();
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
    const size_t total = (size_t)B * N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / N), n = (int)(i - (size_t)b * N);
        const float x0 = pts[((size_t)b * 3 + 0) * N + n], x1 = pts[((size_t)b * 3 + 1) * N + n],
                    x2 = pts[((size_t)b * 3 + 2) * N + n];
        v[0] += x0; v[1] += x1; v[2] += x2;
        v[3] += x0 * x0; v[4] += x0 * x1; v[5] += x0 * x2; v[6] += x1 * x1; v[7] += x1 * x2; v[8] += x2 * x2;
    }
    const float r = warp_reduce_scatter32(v, lane);
    if (lane < 9) atomicAdd(&dred[lane], (double)r);
    __syncthreads();
    if (tid < 9)
        for (int j = 0; j < K; ++j) atomicAdd(&mom[j * GWTF_MOM_STRIDE + tid], dred[tid]);
}

// batch statistics actually used, for the running-stat update and for inspection
// bstat [L][K][2][4][F]: mean0 | var0 | mean1 | var1
static __global__ void k_bstat(gwtf_stack_desc d, const float* params, const double* mom, const double* sum1, double n_total,
                        float* bstat) {
    const int F = d.n_features, K = d.n_components, L = d.n_layers;
    const int total = L * K * 2 * F;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % F, net = (i / F) % 2, j = (i / (2 * F)) % K, l = i / (2 * F * K);
        const unsigned wm = d.warp_mask[l];
        const int w = popc3(wm), k = 3 - w;
        int keepd[3];
        { int q = 0; for (int dd = 0; dd < 3; ++dd) if (!(wm & (1u << dd))) keepd[q++] = dd; }
        const NetOffsets o = net_offsets(F, w);
        const float* P = params + (size_t)(j * L + l) * d.rec_stride + net * o.stride;
        float mean0, var0;
        bn0_from_moments(mom + ((size_t)l * K + j) * GWTF_MOM_STRIDE, n_total, P + o.W0 + c * k, k, keepd, mean0, var0);
        const double* s1 = sum1 + (((size_t)l * K + j) * 2 + net) * 2 * F;
        const double m = s1[c] / n_total;
        const double v = fmax(s1[F + c] / n_total - m * m, 0.0);
        float* o4 = bstat + (((size_t)l * K + j) * 2 + net) * 4 * F;
        o4[c] = mean0; o4[F + c] = var0; o4[2 * F + c] = (float)m; o4[3 * F + c] = (float)v;
    }
}

// per-point mixture NLL from the base-space samples (ubuf slot 0) and the log-det sums
static __global__ void __launch_bounds__(kThreads) k_nll_from_state(int K, int B, int N, const float* __restrict__ z,
                                                             const float* __restrict__ ld, const float* __restrict__ base,
                                                             const float* __restrict__ logw, float* nll, float* logp,
                                                             int32_t* nonfinite) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * N) return;
    const int b = (int)(i / N), n = (int)(i - (size_t)b * N);
    float m = -INFINITY, s = 0.f;
    for (int j = 0; j < K; ++j) {
        float tot = ld[((size_t)j * B + b) * N + n];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float mu = base[b * 6 + d], lv = base[b * 6 + 3 + d];
            const float dz = z[(((size_t)j * B + b) * 3 + d) * N + n] - mu;
            tot += lv + dz * dz / expf(lv);
        }
        const float lp = -0.5f * (tot + 3.0f * GWTF_LOG_2PI);
        if (logp) logp[i * K + j] = lp;
        const float v = lp + logw[b * K + j];
        if (v > m) { s = s * expf(m - v) + 1.0f; m = v; } else s += expf(v - m);
    }
    const float out = -(m + logf(s));
    nll[i] = out;
    if (nonfinite && !isfinite(out)) atomicAdd(nonfinite, 1);      // training.py:43-46: the caller stops on a NaN loss
}

}  // namespace gwtf
