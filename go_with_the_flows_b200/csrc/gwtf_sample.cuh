// Sampling kernel (reference mode='direct', eval-mode BN): flow_mixture.py:141-177 + models.py:199-203.
// A CTA owns a tile of points of one shape: it draws the component index and the base noise of
// every point with Philox4x32-10, compacts the tile by component in shared memory so that a
// warp always runs ONE component's weights (shared-memory broadcast), streams that component's
// layer records with TMA bulk copies and scatters xyz + label back to the point's own slot.
#pragma once
#include "gwtf_common.cuh"

namespace gwtf {

struct SampleArgs {
    gwtf_stack_desc d;
    const float *params, *bnbuf, *film, *base, *cdf;
    int B, N, tiles_per_shape, tile_points;
    uint32_t seed_lo, stream_id;
    const int32_t* idx_in;
    const float* eps_in;
    float* samples;
    int32_t* labels;
    float* z_out;
};

constexpr int kSampleMaxTile = 2048;     // points per CTA tile (multiple of kThreads)
constexpr int kSampleChunk = 4 * kThreads;

template <int FP>
struct SampleSmem {
    LayerW<FP> W;
    uint64_t bar[2];
    int count[GWTF_MAX_COMPONENTS];
    int start[GWTF_MAX_COMPONENTS + 1];
    int cursor[GWTF_MAX_COMPONENTS];
    float zx[kSampleMaxTile][3];              // base-space samples of the tile, in tile order
    unsigned short comp[kSampleMaxTile];
    unsigned short order[kSampleMaxTile];     // tile-local point ids grouped by component
};

// searchsorted(cdf, u, side='right') clamped to K-1
__device__ __forceinline__ int pick_component(const float* cdf, int K, float u) {
    int idx = 0;
    for (int t = 0; t < K; ++t) idx += (cdf[t] <= u) ? 1 : 0;
    return idx < K ? idx : K - 1;
}

// Box-Muller on the kernel's word convention (oracle: box_muller)
__device__ __forceinline__ void draw_normals(uint4 w0, uint4 w1, float (&e)[3]) {
    const float two_pi = 6.283185307179586f;
    const float r01 = sqrtf(-2.0f * logf(u01_open(w0.y)));
    const float th01 = two_pi * u01(w0.z);
    const float r2 = sqrtf(-2.0f * logf(u01_open(w0.w)));
    const float th2 = two_pi * u01(w1.x);
    e[0] = r01 * cosf(th01);
    e[1] = r01 * sinf(th01);
    e[2] = r2 * cosf(th2);
}

struct SampleSched {           // record stream of one CTA: (component, chunk) units x L layers
    int K, L;
    const int* count;
    __device__ __forceinline__ bool next_unit(int& j, int& cb) const {
        if (cb + kSampleChunk < count[j]) { cb += kSampleChunk; return true; }
        ++j; cb = 0;
        while (j < K && count[j] == 0) ++j;
        return j < K;
    }
};

// One (component, chunk) unit: PE points per thread through the L direct layers.
template <int FP, int PE>
__device__ __forceinline__ void sample_chunk(const SampleArgs& a, SampleSmem<FP>& S, float* const (&rawb)[2],
                                             uint32_t (&ph)[2], int& issued, int& consumed, int b, int n0, int j,
                                             int cb, int tid) {
    const int F = a.d.n_features, K = a.d.n_components, L = a.d.n_layers, N = a.N;
    const int cnt = S.count[j] - cb < kSampleChunk ? S.count[j] - cb : kSampleChunk;
    const int st = S.start[j] + cb;
    float x[PE][3];
    int loc[PE];
#pragma unroll
    for (int p = 0; p < PE; ++p) {
        const int i = p * kThreads + tid;
        loc[p] = i < cnt ? S.order[st + i] : -1;
#pragma unroll
        for (int d = 0; d < 3; ++d) x[p][d] = loc[p] >= 0 ? S.zx[loc[p]][d] : 0.f;
    }
    const SampleSched sched{K, L, S.count};
    for (int l = 0; l < L; ++l) {
        const int buf = consumed & 1;
        mbar_wait(&S.bar[buf], ph[buf]);
        ph[buf] ^= 1u;
        ++consumed;
        stage_layer<FP, false>(S.W, nullptr, rawb[buf], LayerSrc(), F, a.d.warp_mask[l], false, false, nullptr, tid,
                               kThreads);
        __syncthreads();
        stage_film<FP, false>(S.W, (LayerWB<FP>*)nullptr, rawb[buf] + rec_stride_of(F) + 8 * F, F, tid, kThreads);
        __syncthreads();
        {   // prefetch the next record of the stream
            int nj = j, ncb = cb, nl = l + 1;
            bool more = true;
            if (nl == L) { nl = 0; more = sched.next_unit(nj, ncb); }
            if (more) {
                if (tid == 0) {
                    LayerSrc src;
                    src.params = a.params + (size_t)(nj * L + nl) * a.d.rec_stride;
                    src.bn = a.bnbuf + (size_t)(nj * L + nl) * 8 * F;
                    src.film = a.film + ((size_t)(b * K + nj) * L + nl) * 4 * F;
                    issue_layer_copy(rawb[issued & 1], src, F, true, true, &S.bar[issued & 1]);
                }
                ++issued;
            }
        }
        if ((tid & ~31) < cnt) {                       // warp-uniform: skip warps with no live point
            float omu[PE][3], olv[PE][3];
            {
                float acc[PE][FP];
                contract_h1<FP, PE>(S.W, 0, F, x, acc);
                head_out<FP, PE>(S.W, 0, acc, omu);
            }
            {
                float acc[PE][FP];
                contract_h1<FP, PE>(S.W, 1, F, x, acc);
                head_out<FP, PE>(S.W, 1, acc, olv);
            }
#pragma unroll
            for (int p = 0; p < PE; ++p) {
                float lam[3];
                warp_point<true>(x[p], omu[p], olv[p], lam);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int p = 0; p < PE; ++p)
        if (loc[p] >= 0)
#pragma unroll
            for (int d = 0; d < 3; ++d) a.samples[((size_t)b * 3 + d) * N + n0 + loc[p]] = x[p][d];
}

template <int FP, int PMAX>
__global__ void __launch_bounds__(kThreads) k_sample(const SampleArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using SM = SampleSmem<FP>;
    SM& S = *reinterpret_cast<SM*>(smem_raw);
    const int F = a.d.n_features, K = a.d.n_components, L = a.d.n_layers;
    const int RAW = round_up(raw_floats(F), 4);
    float* raw0 = reinterpret_cast<float*>(smem_raw + round_up((int)sizeof(SM), 16));
    float* const rawb[2] = {raw0, raw0 + RAW};
    const int tid = threadIdx.x;
    const int b = blockIdx.x / a.tiles_per_shape;
    const int n0 = (blockIdx.x - b * a.tiles_per_shape) * a.tile_points;
    const int N = a.N;

    if (tid == 0) { mbar_init(&S.bar[0], 1); mbar_init(&S.bar[1], 1); mbar_fence_init(); }
    if (tid < GWTF_MAX_COMPONENTS) { S.count[tid] = 0; S.cursor[tid] = 0; }
    __syncthreads();

    // ---- draws
    float mub[3], sdb[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { mub[d] = a.base[b * 6 + d]; sdb[d] = expf(0.5f * a.base[b * 6 + 3 + d]); }
    const float* cdf = a.cdf + (size_t)b * K;
    for (int loc = tid; loc < a.tile_points; loc += kThreads) {
        const int n = n0 + loc;
        int c = 0xFFFF;
        if (n < N) {
            const uint2 key = make_uint2(a.seed_lo, a.stream_id);
            const uint4 w0 = philox4x32_10(make_uint4((uint32_t)n, (uint32_t)b, 0u, 0u), key);
            const uint4 w1 = philox4x32_10(make_uint4((uint32_t)n, (uint32_t)b, 1u, 0u), key);
            c = a.idx_in ? a.idx_in[(size_t)b * N + n] : pick_component(cdf, K, u01(w0.x));
            c = c < 0 ? 0 : (c >= K ? K - 1 : c);
            float e[3];
            if (a.eps_in) {
#pragma unroll
                for (int d = 0; d < 3; ++d) e[d] = a.eps_in[((size_t)b * 3 + d) * N + n];
            } else draw_normals(w0, w1, e);
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float zv = fmaf(sdb[d], e[d], mub[d]);      // models.py:108: eps*std + mu
                S.zx[loc][d] = zv;
                if (a.z_out) a.z_out[((size_t)b * 3 + d) * N + n] = zv;
            }
            a.labels[(size_t)b * N + n] = c + 1;                  // flow_mixture.py:176
            atomicAdd(&S.count[c], 1);
        }
        S.comp[loc] = (unsigned short)c;
    }
    __syncthreads();
    if (tid == 0) {
        int s = 0;
        for (int t = 0; t < K; ++t) { S.start[t] = s; s += S.count[t]; }
        S.start[K] = s;
    }
    __syncthreads();
    for (int loc = tid; loc < a.tile_points; loc += kThreads) {
        const int c = S.comp[loc];
        if (c != 0xFFFF) S.order[S.start[c] + atomicAdd(&S.cursor[c], 1)] = (unsigned short)loc;
    }
    __syncthreads();

    // ---- (component, chunk) units, each L direct layers; records double-buffered by TMA
    uint32_t ph[2] = {0u, 0u};
    int issued = 0, consumed = 0;
    int j = 0, cb = 0;
    while (j < K && S.count[j] == 0) ++j;
    if (j >= K) return;
    if (tid == 0) {
        LayerSrc src;
        src.params = a.params + (size_t)(j * L) * a.d.rec_stride;
        src.bn = a.bnbuf + (size_t)(j * L) * 8 * F;
        src.film = a.film + ((size_t)(b * K + j) * L) * 4 * F;
        issue_layer_copy(rawb[0], src, F, true, true, &S.bar[0]);
    }
    issued = 1;
    const SampleSched sched{K, L, S.count};
    bool more = true;
    while (more) {
        const int rem = S.count[j] - cb;
        if (PMAX >= 4 && rem > 2 * kThreads) sample_chunk<FP, (PMAX >= 4 ? 4 : PMAX)>(a, S, rawb, ph, issued, consumed, b, n0, j, cb, tid);
        else if (PMAX >= 2 && rem > kThreads) sample_chunk<FP, (PMAX >= 2 ? 2 : PMAX)>(a, S, rawb, ph, issued, consumed, b, n0, j, cb, tid);
        else sample_chunk<FP, 1>(a, S, rawb, ph, issued, consumed, b, n0, j, cb, tid);
        more = sched.next_unit(j, cb);
    }
}

}  // namespace gwtf

// =============================================================================================
// Sampling through the per-layer tensor-core kernels.  The points of every shape are regrouped by
// the component they drew inside each shape's row -- (B, 3, Npad), Npad = roundup(N,128) + 128 K: component j's
// points of shape b fill a segment that starts at a multiple of 128 -- pushed through the L direct layers with the
// same kernels the NLL passes use (eval-mode BN, direct=1, segmented tiles), and gathered back to their own slots.
// All sizes are upper bounds known before the draw: no host read-back anywhere.
// =============================================================================================
namespace gwtf {

// Inclusive CDF of softmax(logits) per shape, the way np.random.choice builds it (flow_mixture.py:149-153):
// fp32 probabilities e / sum(e) (sequential fp32 sum), float64 cumsum, normalised, stored fp32 with the last
// entry pinned to 1.  e = fp32(exp(fp64(logit))) -- the correctly rounded fp32 exponential -- so that the
// CPU restatement (oracle/flow_oracle.py: mixture_cdf) and this kernel agree bit for bit.
static __global__ void k_mixture_cdf(const float* __restrict__ logits, int B, int K, float* cdf) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float e[GWTF_MAX_COMPONENTS];
    float sum = 0.f;
    for (int j = 0; j < K; ++j) { e[j] = (float)exp((double)logits[(size_t)b * K + j]); sum += e[j]; }
    double run = 0.0, c[GWTF_MAX_COMPONENTS];
    for (int j = 0; j < K; ++j) { run += (double)(e[j] / sum); c[j] = run; }
    for (int j = 0; j < K; ++j) cdf[(size_t)b * K + j] = j == K - 1 ? 1.0f : (float)(c[j] / run);
}

struct SamplePlanArgs {
    int K, B, N;
    const float* cdf;
    uint32_t seed_lo, stream_id;
    const int32_t* idx_in;
    int32_t* counts;        // (B, K), pre-zeroed
};

__device__ __forceinline__ int sample_component(const float* cdf, int K, const int32_t* idx_in, int b, int n, int N,
                                                uint32_t seed_lo, uint32_t stream_id, uint4& w0) {
    const uint2 key = make_uint2(seed_lo, stream_id);
    w0 = philox4x32_10(make_uint4((uint32_t)n, (uint32_t)b, 0u, 0u), key);
    int c = idx_in ? idx_in[(size_t)b * N + n] : pick_component(cdf, K, u01(w0.x));
    return c < 0 ? 0 : (c >= K ? K - 1 : c);
}

constexpr int kSampleSpan = 16 * kThreads;      // points of one shape a count / scatter CTA covers

// grid (ceil(N / kSampleSpan), B): how many points of each shape drew each component
static __global__ void __launch_bounds__(kThreads) k_sample_count(const SamplePlanArgs a) {
    __shared__ int cnt[GWTF_MAX_COMPONENTS];
    const int b = blockIdx.y, tid = threadIdx.x;
    if (tid < GWTF_MAX_COMPONENTS) cnt[tid] = 0;
    __syncthreads();
    const float* cdf = a.cdf + (size_t)b * a.K;
    const int n1 = min(a.N, (int)(blockIdx.x + 1) * kSampleSpan);
    for (int n = blockIdx.x * kSampleSpan + tid; n < n1; n += kThreads) {
        uint4 w0;
        atomicAdd(&cnt[sample_component(cdf, a.K, a.idx_in, b, n, a.N, a.seed_lo, a.stream_id, w0)], 1);
    }
    __syncthreads();
    if (tid < a.K && cnt[tid]) atomicAdd(&a.counts[b * a.K + tid], cnt[tid]);
}

// one CTA: segment table seg[j][b] = (offset inside shape row b, count), segment starts rounded up to 128 points,
// and seg_tiles[j][0..B] = exclusive prefix over shapes of the segments' tile counts; cursors zeroed.
static __global__ void __launch_bounds__(kThreads) k_sample_plan(int K, int B, const int32_t* counts, int32_t* seg,
                                                                 int32_t* seg_tiles, int32_t* cursor) {
    const int tid = threadIdx.x;
    for (int b = tid; b < B; b += kThreads) {
        int off = 0;
        for (int j = 0; j < K; ++j) {
            const int c = counts[b * K + j];
            seg[((size_t)j * B + b) * 2] = off;
            seg[((size_t)j * B + b) * 2 + 1] = c;
            cursor[b * K + j] = 0;
            off += (c + 127) / 128 * 128;
        }
    }
    __syncthreads();
    if (tid < K) {
        int run = 0;
        for (int b = 0; b < B; ++b) {
            seg_tiles[(size_t)tid * (B + 1) + b] = run;
            run += (counts[b * K + tid] + 127) / 128;
        }
        seg_tiles[(size_t)tid * (B + 1) + B] = run;
    }
}

struct SampleScatterArgs {
    int K, B, N, Npad;
    const float *base, *cdf;
    uint32_t seed_lo, stream_id;
    const int32_t* idx_in;
    const float* eps_in;
    const int32_t* seg;     // (K, B, 2)
    int32_t* cursor;        // (B, K) running fill of every segment
    float* xbuf;            // (B, 3, Npad) base-space samples, points grouped by component inside each row
    int32_t* slot;          // (B, N): position of the point inside its row
    int32_t* labels;
    float* z_out;
};

// grid (ceil(N / kSampleSpan), B): draw, reserve a range of each component's segment for this CTA (one global atomic
// per component), place the points (order inside a segment is free: results are gathered back by slot)
static __global__ void __launch_bounds__(kThreads) k_sample_scatter(const SampleScatterArgs a) {
    __shared__ int cnt[GWTF_MAX_COMPONENTS], start[GWTF_MAX_COMPONENTS];
    const int b = blockIdx.y, tid = threadIdx.x, N = a.N, K = a.K;
    if (tid < GWTF_MAX_COMPONENTS) cnt[tid] = 0;
    __syncthreads();
    float mub[3], sdb[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { mub[d] = a.base[b * 6 + d]; sdb[d] = expf(0.5f * a.base[b * 6 + 3 + d]); }
    const float* cdf = a.cdf + (size_t)b * K;
    constexpr int PER = kSampleSpan / kThreads;
    int comp[PER], rank[PER];
    const int nb = blockIdx.x * kSampleSpan;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int n = nb + i * kThreads + tid;
        comp[i] = -1;
        if (n < N) {
            uint4 w0;
            comp[i] = sample_component(cdf, K, a.idx_in, b, n, N, a.seed_lo, a.stream_id, w0);
            rank[i] = atomicAdd(&cnt[comp[i]], 1);
        }
    }
    __syncthreads();
    if (tid < K) start[tid] = cnt[tid] ? atomicAdd(&a.cursor[b * K + tid], cnt[tid]) : 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int n = nb + i * kThreads + tid;
        if (comp[i] < 0) continue;
        const int c = comp[i];
        float e[3];
        if (a.eps_in) {
#pragma unroll
            for (int d = 0; d < 3; ++d) e[d] = a.eps_in[((size_t)b * 3 + d) * N + n];
        } else {
            const uint2 key = make_uint2(a.seed_lo, a.stream_id);
            const uint4 w0 = philox4x32_10(make_uint4((uint32_t)n, (uint32_t)b, 0u, 0u), key);
            const uint4 w1 = philox4x32_10(make_uint4((uint32_t)n, (uint32_t)b, 1u, 0u), key);
            draw_normals(w0, w1, e);
        }
        const int pos = a.seg[((size_t)c * a.B + b) * 2] + start[c] + rank[i];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float zv = fmaf(sdb[d], e[d], mub[d]);          // models.py:108: eps*std + mu
            a.xbuf[((size_t)b * 3 + d) * a.Npad + pos] = zv;
            if (a.z_out) a.z_out[((size_t)b * 3 + d) * N + n] = zv;
        }
        a.slot[(size_t)b * N + n] = pos;
        a.labels[(size_t)b * N + n] = c + 1;                      // flow_mixture.py:176
    }
}

static __global__ void __launch_bounds__(kThreads) k_sample_gather(int B, int N, int Npad, const float* xbuf,
                                                                   const int32_t* slot, float* samples) {
    const size_t total = (size_t)B * N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / N), n = (int)(i - (size_t)b * N);
        const int pos = slot[i];
#pragma unroll
        for (int d = 0; d < 3; ++d)
            samples[((size_t)b * 3 + d) * N + n] = xbuf[((size_t)b * 3 + d) * Npad + pos];
    }
}

}  // namespace gwtf
