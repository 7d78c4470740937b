// C ABI of libgwtf.so (see include/gwtf.h).  Host-side argument checks, engine dispatch, the single-call
// drivers, the rank exchange, the fused AMSGrad step and the roofline probes.  Launches go on the caller's
// stream; nothing here allocates device memory or synchronises the host (the two probes excepted), and there
// is no mutable process-global state: options ride in the caller's gwtf_stack_desc.
#include <new>

#include "gwtf_host.h"
#include "gwtf_fwd.cuh"
#include "gwtf_bwd.cuh"
#include "gwtf_sample.cuh"
#include "gwtf_exchange.cuh"

using namespace gwtf;

// peer-memory statistic exchange of one rank (gwtf_exchange_create)
struct gwtf_exchange {
    int rank = 0, world = 1, slot = 0;
    unsigned long long seq = 0;
    unsigned long long timeout_ns = 600ull * 1000000000ull;
    double* recv[kMaxRanks] = {};
    unsigned long long* flags[kMaxRanks] = {};
    int ll = 1;                 // GWTF_EXCHANGE_LL=0: data -> fence -> flag protocol
};

namespace gwtf {

char* err_buf() {
    static thread_local char buf[512] = "";
    return buf;
}

size_t keep_layer_floats(const gwtf_stack_desc& d, int B, int N) {
    const int F = d.n_features, K = d.n_components;
    switch (bwd_engine(d)) {
        case kEngineTc: return 0;                                        // the tcgen05 backward recomputes
        case kEngineMma: return (size_t)K * mma_keep_floats(F, B, N);
        default: return (size_t)K * 2 * F * B * N;
    }
}

}  // namespace gwtf

namespace {

int check_desc(const gwtf_stack_desc* d) {
    if (!d) return fail(-1, "null stack descriptor");
    if (d->n_components < 1 || d->n_components > GWTF_MAX_COMPONENTS) return fail(-2, "n_components out of range");
    if (d->n_layers < 1 || d->n_layers > GWTF_MAX_LAYERS) return fail(-3, "n_layers out of range");
    if (d->n_features < 1 || d->n_features > GWTF_MAX_FEATURES) return fail(-4, "n_features out of range (1..64)");
    if (d->rec_stride != rec_stride_of(d->n_features)) return fail(-5, "rec_stride does not match gwtf_rec_stride(F)");
    for (int l = 0; l < d->n_layers; ++l) {
        const int w = __builtin_popcount(d->warp_mask[l] & 7);
        if (w < 1 || w > 2 || (d->warp_mask[l] & ~7)) return fail(-6, "warp_mask must select 1 or 2 of the 3 dims");
    }
    const int e = d->engine;
    if (!(e == GWTF_ENGINE_DEFAULT || e == GWTF_ENGINE_FMA || e == GWTF_ENGINE_TC_FWD || e == GWTF_ENGINE_MMA ||
          e == GWTF_ENGINE_TC))
        return fail(-7, "unknown engine");
    return 0;
}

int exchange_sum(gwtf_exchange* x, double* data, int n, bool pdl, cudaStream_t st) {
    if (!x || x->world <= 1) return 0;
    if (n > x->slot) return fail(-20, "exchange slot too small for this stack");
    ExchangeArgs a;
    a.rank = x->rank; a.world = x->world; a.n = n; a.slot = x->slot; a.seq = ++x->seq; a.data = data;
    a.timeout_ns = x->timeout_ns; a.ll = x->ll;
    for (int r = 0; r < kMaxRanks; ++r) { a.recv[r] = x->recv[r]; a.flags[r] = x->flags[r]; }
    GWTF_CUDA(launch_pdl(pdl, k_exchange_sum, dim3(1), dim3(n > 512 ? 1024 : 256), 0, st, a));
    return 0;
}

// the same exchange as the tail of the kernel that completes `data` (tcgen05 layer kernels; gwtf_exchange.cuh)
ExchangeTail make_tail(gwtf_exchange* x, double* data, int n) {
    ExchangeTail t;
    t.x.rank = x->rank; t.x.world = x->world; t.x.n = n; t.x.slot = x->slot; t.x.seq = ++x->seq; t.x.data = data;
    t.x.timeout_ns = x->timeout_ns; t.x.ll = x->ll;
    for (int r = 0; r < kMaxRanks; ++r) { t.x.recv[r] = x->recv[r]; t.x.flags[r] = x->flags[r]; }
    return t;
}
ExchangeTail no_tail() { return ExchangeTail(); }
// GWTF_EXCHANGE_FOLD=1: run the exchange in the tail of the producing tcgen05 kernel instead of a kernel of its own
// (bring-up switch; measured gain on 2 GPUs: 0.1 ms of a 13.9 ms step)
bool fold_exchange() {
    static const bool on = [] { const char* e = getenv("GWTF_EXCHANGE_FOLD"); return e && e[0] == '1'; }();
    return on;
}

int fwd_layer_dispatch(const LayerArgs& a, int phase, cudaStream_t st) {
    switch (fwd_engine(a.d)) {
        case kEngineFma: return a.seg ? fail(-4, "segmented rows need the tcgen05 forward") : launch_fwd_layer_fma(a, phase, st);
        case kEngineMma: return launch_fwd_layer_mma(a, phase, st);
        default: return launch_fwd_layer_tc(a, phase, st);
    }
}

}  // namespace

namespace {
__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, float* __restrict__ vmax, long long n, float lr,
                                              float beta1, float beta2, float eps, float wd, float bc1, float bc2) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gi = g[i];
        const float mi = m[i] * beta1 + (1.0f - beta1) * gi;
        float vi = v[i] * beta2;
        vi = vi + (1.0f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        float den = vi;
        if (vmax) { den = fmaxf(vmax[i], vi); vmax[i] = den; }
        const float denom = sqrtf(den) / bc2 + eps;
        const float mc = mi / bc1;
        const float pi = p[i];
        p[i] = wd != 0.0f ? pi - (pi * wd + lr * (mc / denom)) : pi - lr * (mc / denom);
    }
}
}  // namespace

namespace {
// issue-rate probe of the warp-level tensor-core path: 16 independent m16n8k8 tf32 accumulators per warp
__global__ void __launch_bounds__(256) k_mma_probe(float* out, int iters) {
    float d[16][4];
    uint32_t a[4], bq[2];
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
    for (int i = 0; i < 2; ++i) bq[i] = __float_as_uint(0.5f + i);
#pragma unroll
    for (int n = 0; n < 16; ++n) d[n][0] = d[n][1] = d[n][2] = d[n][3] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int n = 0; n < 16; ++n) mma_tf32(d[n], a, bq[0], bq[1]);
    }
    float s = 0.f;
#pragma unroll
    for (int n = 0; n < 16; ++n) s += d[n][0] + d[n][1] + d[n][2] + d[n][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_ffma_probe(float* out, int iters, float a, float b) {
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename Launch>
int time_probe(Launch launch, int grid, float& best_ms, cudaStream_t st) {
    float* out = nullptr;
    GWTF_CUDA(cudaMalloc(&out, sizeof(float) * grid * 256));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(out);
    best_ms = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0, st);
        launch(out);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaError_t err = cudaGetLastError();
    cudaFree(out);
    if (err != cudaSuccess) return cuda_fail(err, "roofline probe");
    return 0;
}
}  // namespace

extern "C" {

int gwtf_version(void) { return 2; }
const char* gwtf_last_error_string(void) { return err_buf(); }

int gwtf_resolved_engine(const gwtf_stack_desc* desc, int32_t which) {
    if (check_desc(desc)) return -1;
    return which == 0 ? fwd_engine(*desc) : bwd_engine(*desc);
}

int64_t gwtf_keep_floats(const gwtf_stack_desc* desc, int32_t B, int32_t N) {
    if (check_desc(desc) || B <= 0 || N <= 0) return 0;
    return (int64_t)desc->n_layers * (int64_t)keep_layer_floats(*desc, B, N);
}

int gwtf_rec_stride(int32_t F) { return rec_stride_of(F); }

int gwtf_param_offsets(int32_t F, int32_t n_warp, int32_t* offsets, int32_t* net_stride) {
    if (!offsets || !net_stride || n_warp < 1 || n_warp > 2) return fail(-1, "bad arguments to gwtf_param_offsets");
    const NetOffsets o = net_offsets(F, n_warp);
    offsets[0] = o.W0; offsets[1] = o.g0; offsets[2] = o.b0; offsets[3] = o.W1; offsets[4] = o.W2; offsets[5] = o.b2;
    *net_stride = o.stride;
    return 0;
}

int gwtf_nll_fwd_eval(const gwtf_stack_desc* desc, const float* params, const float* bnbuf, const float* film,
                      const float* points, const float* base, const float* logw, int32_t B, int32_t N, float* nll,
                      float* logp, float* z, float* ssum, void* stream) {
    if (int rc = check_desc(desc)) return rc;
    if (!params || !bnbuf || !film || !points || !base || !logw || !nll) return fail(-10, "null pointer argument");
    if (B < 0 || N < 0) return fail(-11, "negative size");
    if (B == 0 || N == 0) return 0;
    EvalArgs a;
    a.d = *desc; a.params = params; a.bnbuf = bnbuf; a.film = film; a.points = points; a.base = base; a.logw = logw;
    a.B = B; a.N = N; a.nll = nll; a.logp = logp; a.z = z; a.ssum = ssum; a.tiles_per_shape = 0;
    return launch_eval_fma(a, (cudaStream_t)stream);
}

int gwtf_fwd_moments(const gwtf_stack_desc* desc, const float* points, int32_t B, int32_t N, double* mom, void* stream) {
    if (int rc = check_desc(desc)) return rc;
    if (!points || !mom) return fail(-10, "null pointer argument");
    if (B <= 0 || N <= 0) return 0;
    return launch_moments(points, B, N, desc->n_components, mom, (cudaStream_t)stream);
}

// NOTE: the C ABI takes explicit per-layer pointers so a multi-rank caller can all-reduce
// mom / sum1 between phases; xin_shared=1 when `xin` is the (B,3,N) data cloud.
static int fwd_layer_ex_impl(const gwtf_stack_desc* desc, int32_t layer, int32_t phase, int32_t train, int32_t direct,
                             const float* params, const float* bnbuf, const float* film, const float* xin,
                             int32_t xin_shared, float* xout, float* ld, float* ssum, float* trio, float* y1out,
                             const double* mom_in, double* mom_out, double* sum1, int32_t B, int32_t N, double n_total,
                             const ExchangeTail* tail, void* stream) {
    if (int rc = check_desc(desc)) return rc;
    if (layer < 0 || layer >= desc->n_layers) return fail(-12, "layer out of range");
    if (phase != 0 && phase != 1) return fail(-13, "phase must be 0 or 1");
    if (!params || !xin) return fail(-10, "null pointer argument");
    if (train && (!mom_in || !sum1)) return fail(-10, "train mode needs mom_in and sum1");
    if (!train && !bnbuf) return fail(-10, "eval mode needs bnbuf");
    if (phase == 1 && (!film || !xout)) return fail(-10, "phase 1 needs film and xout");
    if (B <= 0 || N <= 0) return 0;
    LayerArgs a;
    a.d = *desc; a.layer = layer; a.train = train; a.direct = direct; a.params = params; a.bnbuf = bnbuf; a.film = film;
    a.xin = xin; a.xin_shared = xin_shared; a.xout = xout; a.ld = ld; a.ssum = ssum; a.trio = trio; a.y1out = y1out;
    a.mom_in = mom_in; a.mom_out = mom_out; a.sum1 = sum1; a.B = B; a.N = N; a.n_total = n_total; a.tiles_per_shape = 0;
    a.seg = nullptr; a.seg_tiles = nullptr;
    a.tail = tail ? *tail : no_tail();
    return fwd_layer_dispatch(a, phase, (cudaStream_t)stream);
}

int gwtf_fwd_layer_ex(const gwtf_stack_desc* desc, int32_t layer, int32_t phase, int32_t train, int32_t direct,
                      const float* params, const float* bnbuf, const float* film, const float* xin,
                      int32_t xin_shared, float* xout, float* ld, float* ssum, float* trio, float* y1out,
                      const double* mom_in, double* mom_out, double* sum1, int32_t B, int32_t N, double n_total,
                      void* stream) {
    return fwd_layer_ex_impl(desc, layer, phase, train, direct, params, bnbuf, film, xin, xin_shared, xout, ld, ssum, trio,
                             y1out, mom_in, mom_out, sum1, B, N, n_total, nullptr, stream);
}

static int fwd_layer_impl(const gwtf_stack_desc* desc, int32_t layer, int32_t phase, int32_t train, const float* params,
                          const float* bnbuf, const float* film, const float* points, float* ubuf, float* ld, float* ssum,
                          float* ybuf, double* mom, double* sum1, int32_t B, int32_t N, double n_total,
                          const ExchangeTail* tail, void* stream) {
    if (int rc = check_desc(desc)) return rc;
    if (!ubuf) return fail(-10, "null pointer argument");
    const int L = desc->n_layers, K = desc->n_components, F = desc->n_features;
    if (layer < 0 || layer >= L) return fail(-12, "layer out of range");
    const size_t slot = (size_t)K * B * 3 * N;
    const bool first = layer == L - 1;
    const float* xin = first ? points : ubuf + (size_t)(layer + 1) * slot;
    double* mom_in = mom ? mom + (size_t)layer * K * GWTF_MOM_STRIDE : nullptr;
    double* mom_out = (mom && layer > 0) ? mom + (size_t)(layer - 1) * K * GWTF_MOM_STRIDE : nullptr;
    double* s1 = sum1 ? sum1 + (size_t)layer * K * 4 * F : nullptr;
    float* y1 = ybuf ? ybuf + (size_t)layer * keep_layer_floats(*desc, B, N) : nullptr;
    return fwd_layer_ex_impl(desc, layer, phase, train, 0, params, bnbuf, film, xin, first ? 1 : 0,
                             ubuf + (size_t)layer * slot, ld, ssum, nullptr, y1, mom_in, train ? mom_out : nullptr, s1,
                             B, N, n_total, tail, stream);
}

int gwtf_fwd_layer(const gwtf_stack_desc* desc, int32_t layer, int32_t phase, int32_t train, const float* params,
                   const float* bnbuf, const float* film, const float* points, float* ubuf, float* ld, float* ssum,
                   float* ybuf, double* mom, double* sum1, int32_t B, int32_t N, double n_total, void* stream) {
    return fwd_layer_impl(desc, layer, phase, train, params, bnbuf, film, points, ubuf, ld, ssum, ybuf, mom, sum1, B, N,
                          n_total, nullptr, stream);
}

int gwtf_fwd_bstat(const gwtf_stack_desc* desc, const float* params, const double* mom, const double* sum1,
                   double n_total, float* bstat, void* stream) {
    if (int rc = check_desc(desc)) return rc;
    if (!params || !mom || !sum1 || !bstat) return fail(-10, "null pointer argument");
    return launch_bstat(*desc, params, mom, sum1, n_total, bstat, (cudaStream_t)stream);
}

int gwtf_nll_from_state(const gwtf_stack_desc* desc, const float* ubuf, const float* ld, const float* base,
                        const float* logw, int32_t B, int32_t N, float* nll, float* logp, void* stream) {
    if (int rc = check_desc(desc)) return rc;
    if (!ubuf || !ld || !base || !logw || !nll) return fail(-10, "null pointer argument");
    if (B <= 0 || N <= 0) return 0;
    return launch_nll_from_state(*desc, B, N, ubuf, ld, base, logw, nll, logp, (cudaStream_t)stream);
}

int64_t gwtf_eval_layers_workspace_bytes(const gwtf_stack_desc* desc, int32_t B, int32_t N) {
    if (check_desc(desc) || B <= 0 || N <= 0) return 0;
    const int64_t K = desc->n_components;
    return 4 * (2 * K * B * 3 * N + K * B * N);              // two ping-pong coordinate slots + the log-det sums
}

// Eval-mode NLL through the per-layer kernels of the descriptor's engine (tensor cores): L launches that
// ping-pong between two (K,B,3,N) slots of the workspace, then the mixture head.  Faster than the single-launch
// FMA kernel at every size measured on B200.  Honours desc->eval_precision.
int gwtf_nll_fwd_eval_layers(const gwtf_stack_desc* desc, const float* params, const float* bnbuf, const float* film,
                             const float* points, const float* base, const float* logw, void* workspace,
                             int64_t workspace_bytes, int32_t B, int32_t N, float* nll, float* logp, void* stream) {
    if (int rc = check_desc(desc)) return rc;
    if (!params || !bnbuf || !film || !points || !base || !logw || !workspace || !nll)
        return fail(-10, "null pointer argument");
    if (B <= 0 || N <= 0) return 0;
    if (workspace_bytes < gwtf_eval_layers_workspace_bytes(desc, B, N)) return fail(-25, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int L = desc->n_layers, K = desc->n_components;
    const size_t slot = (size_t)K * B * 3 * N;
    float* scratch = (float*)workspace;
    float* ld = scratch + 2 * slot;
    GWTF_CUDA(cudaMemsetAsync(ld, 0, sizeof(float) * (size_t)K * B * N, st));
    const float* xin = points;
    for (int l = L - 1; l >= 0; --l) {
        float* xout = scratch + (size_t)(l & 1) * slot;
        if (int rc = gwtf_fwd_layer_ex(desc, l, 1, 0, 0, params, bnbuf, film, xin, l == L - 1 ? 1 : 0, xout, ld, nullptr,
                                       nullptr, nullptr, nullptr, nullptr, nullptr, B, N, (double)B * (double)N, stream))
            return rc;
        xin = xout;
    }
    return gwtf_nll_from_state(desc, scratch, ld, base, logw, B, N, nll, logp, stream);
}

int gwtf_exchange_create(int32_t rank, int32_t world, void* const* recv, void* const* flags, int32_t slot_doubles,
                         double timeout_s, gwtf_exchange** out) {
    if (!out) return fail(-10, "null pointer argument");
    *out = nullptr;
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) return fail(-21, "bad rank / world size");
    if (world > 1 && (!recv || !flags || slot_doubles <= 0)) return fail(-21, "exchange buffers missing");
    for (int r = 0; r < world && world > 1; ++r)
        if (!recv[r] || !flags[r]) return fail(-21, "null peer pointer");
    gwtf_exchange* x = new (std::nothrow) gwtf_exchange();
    if (!x) return fail(-26, "out of host memory");
    x->rank = rank; x->world = world; x->slot = slot_doubles;
    if (const char* e = getenv("GWTF_EXCHANGE_LL")) x->ll = e[0] != '0';
    if (timeout_s > 0.0) x->timeout_ns = (unsigned long long)(timeout_s * 1e9);
    for (int r = 0; r < world && world > 1; ++r) {
        x->recv[r] = (double*)recv[r];
        x->flags[r] = (unsigned long long*)flags[r];
    }
    *out = x;
    return 0;
}
int gwtf_exchange_destroy(gwtf_exchange* x) { delete x; return 0; }
int gwtf_exchange_world(const gwtf_exchange* x) { return x ? x->world : 1; }
uint64_t gwtf_exchange_seq(const gwtf_exchange* x) { return x ? x->seq : 0; }
int gwtf_exchange_resync(gwtf_exchange* x, uint64_t seq) {
    if (!x) return fail(-10, "null pointer argument");
    if (seq < x->seq) return fail(-27, "the sequence number may only move forward (peers wait for >= seq)");
    x->seq = seq;
    return 0;
}
int gwtf_exchange_sum(gwtf_exchange* x, double* data, int32_t n, void* stream) {
    if (!data || n <= 0) return fail(-10, "null pointer argument");
    return exchange_sum(x, data, n, true, (cudaStream_t)stream);
}

// n_total > 0: statistics are over n_total points on all ranks; the sums are exchanged between phases
// through desc->exchange
static int fwd_all_impl(const gwtf_stack_desc* desc, int32_t train, const float* params, const float* bnbuf,
                        const float* film, const float* points, const float* base, const float* logw, float* ubuf,
                        float* ld, float* ssum, float* ybuf, double* mom, double* sum1, float* bstat, int32_t B,
                        int32_t N, float* nll, float* logp, double n_total_in, void* stream) {
    if (int rc = check_desc(desc)) return rc;
    if (!ubuf || !ld) return fail(-10, "null pointer argument");
    if (B <= 0 || N <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int L = desc->n_layers, K = desc->n_components, F = desc->n_features;
    const bool ranks = n_total_in > 0.0 && train;
    gwtf_exchange* x = desc->exchange;
    if (ranks && (!x || x->world <= 1)) return fail(-22, "the descriptor carries no rank exchange (gwtf_exchange_create)");
    const bool pdl = pdl_on(*desc);
    const double n_total = n_total_in > 0.0 ? n_total_in : (double)B * (double)N;
    GWTF_CUDA(cudaMemsetAsync(ld, 0, sizeof(float) * (size_t)K * B * N, st));
    if (ssum) GWTF_CUDA(cudaMemsetAsync(ssum, 0, sizeof(float) * (size_t)K * B * 3 * N, st));
    if (train) {
        if (!mom || !sum1) return fail(-10, "train mode needs mom and sum1");
        GWTF_CUDA(cudaMemsetAsync(mom, 0, sizeof(double) * (size_t)L * K * GWTF_MOM_STRIDE, st));
        GWTF_CUDA(cudaMemsetAsync(sum1, 0, sizeof(double) * (size_t)L * K * 4 * F, st));
        if (int rc = gwtf_fwd_moments(desc, points, B, N, mom + (size_t)(L - 1) * K * GWTF_MOM_STRIDE, stream)) return rc;
    }
    // tcgen05 kernels run the exchange of the sums they complete in their own tail (the statistics pass: sum1 of its
    // layer, the apply pass: the moments of the next layer); the other engines get the stand-alone exchange kernel
    const bool fold = ranks && fwd_engine(*desc) == kEngineTcFwd && fold_exchange();
    if (fold && K * 4 * F > x->slot) return fail(-20, "exchange slot too small for this stack");
    for (int l = L - 1; l >= 0; --l) {
        if (train) {
            if (ranks && (!fold || l == L - 1))
                if (int rc = exchange_sum(x, mom + (size_t)l * K * GWTF_MOM_STRIDE, K * GWTF_MOM_STRIDE, pdl, st)) return rc;
            ExchangeTail t0 = fold ? make_tail(x, sum1 + (size_t)l * K * 4 * F, K * 4 * F) : no_tail();
            if (int rc = fwd_layer_impl(desc, l, 0, 1, params, bnbuf, film, points, ubuf, ld, ssum, ybuf, mom, sum1, B,
                                        N, n_total, &t0, stream)) return rc;
            if (ranks && !fold) if (int rc = exchange_sum(x, sum1 + (size_t)l * K * 4 * F, K * 4 * F, pdl, st)) return rc;
        }
        ExchangeTail t1 = (fold && l > 0) ? make_tail(x, mom + (size_t)(l - 1) * K * GWTF_MOM_STRIDE, K * GWTF_MOM_STRIDE)
                                          : no_tail();
        if (int rc = fwd_layer_impl(desc, l, 1, train, params, bnbuf, film, points, ubuf, ld, ssum, ybuf, mom, sum1, B,
                                    N, n_total, &t1, stream)) return rc;
    }
    if (train && bstat)
        if (int rc = gwtf_fwd_bstat(desc, params, mom, sum1, n_total, bstat, stream)) return rc;
    if (nll) return gwtf_nll_from_state(desc, ubuf, ld, base, logw, B, N, nll, logp, stream);
    return 0;
}

int gwtf_fwd_all(const gwtf_stack_desc* desc, int32_t train, const float* params, const float* bnbuf,
                 const float* film, const float* points, const float* base, const float* logw, float* ubuf, float* ld,
                 float* ssum, float* ybuf, double* mom, double* sum1, float* bstat, int32_t B, int32_t N, float* nll,
                 float* logp, void* stream) {
    return fwd_all_impl(desc, train, params, bnbuf, film, points, base, logw, ubuf, ld, ssum, ybuf, mom, sum1, bstat, B, N,
                        nll, logp, 0.0, stream);
}
int gwtf_fwd_all_ranks(const gwtf_stack_desc* desc, int32_t train, const float* params, const float* bnbuf,
                       const float* film, const float* points, const float* base, const float* logw, float* ubuf,
                       float* ld, float* ssum, float* ybuf, double* mom, double* sum1, float* bstat, int32_t B,
                       int32_t N, float* nll, float* logp, double n_total, void* stream) {
    if (!(n_total > 0.0)) return fail(-23, "n_total must be the number of points on all ranks");
    return fwd_all_impl(desc, train, params, bnbuf, film, points, base, logw, ubuf, ld, ssum, ybuf, mom, sum1, bstat, B, N,
                        nll, logp, n_total, stream);
}

// ------------------------------------------------------------------------------------------ backward
int gwtf_bwd_seed(const gwtf_stack_desc* desc, const float* ubuf, const float* ld, const float* base,
                  const float* logw, const float* nll, const float* dnll, int32_t B, int32_t N, float* gbuf, float* gs,
                  float* dbase, float* dlogw, void* stream) {
    if (int rc = check_desc(desc)) return rc;
    if (!ubuf || !ld || !base || !logw || !nll || !dnll || !gbuf || !gs || !dbase || !dlogw)
        return fail(-10, "null pointer argument");
    if (B <= 0 || N <= 0) return 0;
    return launch_bwd_seed(desc->n_components, B, N, ubuf, ld, base, logw, nll, dnll, gbuf, gs, dbase, dlogw,
                           (cudaStream_t)stream);
}

static int bwd_layer_impl(const gwtf_stack_desc* desc, int32_t layer, int32_t phase, int32_t train, const float* params,
                          const float* bnbuf, const float* film, const float* points, const float* ubuf, const float* ybuf,
                          const double* mom, const double* sum1, double* bsum, float* gbuf, const float* gs, float* dobuf,
                          float* dparams, float* dfilm, int32_t B, int32_t N, double n_total, const ExchangeTail* tail,
                          void* stream) {
    if (int rc = check_desc(desc)) return rc;
    const int L = desc->n_layers, K = desc->n_components, F = desc->n_features;
    if (layer < 0 || layer >= L) return fail(-12, "layer out of range");
    if (phase != 0 && phase != 1) return fail(-13, "phase must be 0 or 1");
    if (!params || !film || !points || !ubuf || !bsum || !gbuf || !gs || !dobuf || !dparams || !dfilm)
        return fail(-10, "null pointer argument");
    if (train && (!mom || !sum1)) return fail(-10, "train mode needs mom and sum1");
    if (!train && !bnbuf) return fail(-10, "eval mode needs bnbuf");
    if (B <= 0 || N <= 0) return 0;
    const size_t slot = (size_t)K * B * 3 * N;
    const bool first = layer == L - 1;
    BwdArgs a;
    a.d = *desc; a.layer = layer; a.train = train; a.params = params; a.bnbuf = bnbuf; a.film = film;
    a.xin = first ? points : ubuf + (size_t)(layer + 1) * slot;
    a.xin_shared = first ? 1 : 0;
    a.xout = ubuf + (size_t)layer * slot;
    a.y1in = ybuf ? ybuf + (size_t)layer * keep_layer_floats(*desc, B, N) : nullptr;
    a.kept_y1 = (a.y1in && fwd_engine(*desc) == kEngineTcFwd) ? 1 : 0;
    a.mom_in = mom ? mom + (size_t)layer * K * GWTF_MOM_STRIDE : nullptr;
    a.sum1 = sum1 ? sum1 + (size_t)layer * K * 4 * F : nullptr;
    a.bsum = bsum + (size_t)layer * K * 8 * F;
    a.mom_prev = (train && layer > 0) ? mom + (size_t)(layer - 1) * K * GWTF_MOM_STRIDE : nullptr;
    a.bsum_prev = (train && layer > 0) ? bsum + (size_t)(layer - 1) * K * 8 * F : nullptr;
    a.gbuf = gbuf; a.gs = gs; a.dobuf = dobuf; a.dparams = dparams; a.dfilm = dfilm;
    a.B = B; a.N = N; a.n_total = n_total; a.tiles_per_shape = 0;
    a.tail = tail ? *tail : no_tail();
    cudaStream_t st = (cudaStream_t)stream;
    switch (bwd_engine(*desc)) {
        case kEngineFma: return launch_bwd_layer_fma(a, phase, st);
        case kEngineTc: a.y1in = nullptr; a.kept_y1 = 0; return launch_bwd_layer_tc(a, phase, st);
        default: return launch_bwd_layer_mma(a, phase, st);
    }
}

int gwtf_bwd_layer(const gwtf_stack_desc* desc, int32_t layer, int32_t phase, int32_t train, const float* params,
                   const float* bnbuf, const float* film, const float* points, const float* ubuf, const float* ybuf,
                   const double* mom, const double* sum1, double* bsum, float* gbuf, const float* gs, float* dobuf,
                   float* dparams, float* dfilm, int32_t B, int32_t N, double n_total, void* stream) {
    return bwd_layer_impl(desc, layer, phase, train, params, bnbuf, film, points, ubuf, ybuf, mom, sum1, bsum, gbuf, gs,
                          dobuf, dparams, dfilm, B, N, n_total, nullptr, stream);
}

int gwtf_bwd_finish(const gwtf_stack_desc* desc, int32_t train, const float* params, const float* bnbuf,
                    const double* mom, const double* bsum, const float* gbuf, const float* points, float* dparams,
                    float* dpoints, int32_t B, int32_t N, double n_total, void* stream) {
    if (int rc = check_desc(desc)) return rc;
    if (!params || !dparams) return fail(-10, "null pointer argument");
    if (train && (!mom || !bsum)) return fail(-10, "train mode needs mom and bsum");
    if (!train && !bnbuf) return fail(-10, "eval mode needs bnbuf");
    if (B <= 0 || N <= 0) return 0;
    if (dpoints && (!gbuf || !points)) return fail(-10, "dpoints needs gbuf and points");
    FinishArgs fa;
    fa.d = *desc; fa.train = train; fa.params = params; fa.bnbuf = bnbuf; fa.mom = mom; fa.bsum = bsum;
    fa.dparams = dparams; fa.n_total = n_total; fa.local_frac = ((double)B * (double)N) / n_total;
    return launch_bwd_finish(fa, *desc, train, params, mom, bsum, gbuf, points, dpoints, B, N, n_total,
                             (cudaStream_t)stream);
}

static int bwd_all_impl(const gwtf_stack_desc* desc, int32_t train, const float* params, const float* bnbuf,
                        const float* film, const float* points, const float* base, const float* logw, const float* ubuf,
                        const float* ybuf, const float* ld, const double* mom, const double* sum1, const float* nll,
                        const float* dnll, double* bsum, float* gbuf, float* gs, float* dobuf, float* dparams,
                        float* dfilm, float* dbase, float* dlogw, float* dpoints, int32_t B, int32_t N,
                        double n_total_in, void* stream) {
    if (int rc = check_desc(desc)) return rc;
    if (B <= 0 || N <= 0) return 0;
    const bool ranks = n_total_in > 0.0 && train;
    gwtf_exchange* x = desc->exchange;
    if (ranks && (!x || x->world <= 1)) return fail(-22, "the descriptor carries no rank exchange (gwtf_exchange_create)");
    const double n_total = n_total_in > 0.0 ? n_total_in : (double)B * (double)N;
    const int K = desc->n_components, F = desc->n_features;
    if (dnll) {   // seeds from the in-kernel NLL; otherwise gbuf / gs already hold dL/dz, dL/dS
        if (int rc = gwtf_bwd_seed(desc, ubuf, ld, base, logw, nll, dnll, B, N, gbuf, gs, dbase, dlogw, stream)) return rc;
    }
    const bool fold = ranks && bwd_engine(*desc) == kEngineTc && fold_exchange();   // the tcgen05 kernels exchange in their own tail
    if (fold && K * 8 * F > x->slot) return fail(-20, "exchange slot too small for this stack");
    for (int l = 0; l < desc->n_layers; ++l)
        for (int phase = 0; phase < 2; ++phase) {
            // phase 0 completes the sd1_bn sums (slots 0,1), phase 1 the bn0 sums (slots 2,3); the slots of
            // the other phase ride along (nobody reads slots 0,1 after phase 1)
            ExchangeTail t = fold ? make_tail(x, bsum + (size_t)l * K * 8 * F, K * 8 * F) : no_tail();
            if (int rc = bwd_layer_impl(desc, l, phase, train, params, bnbuf, film, points, ubuf, ybuf, mom, sum1, bsum,
                                        gbuf, gs, dobuf, dparams, dfilm, B, N, n_total, &t, stream)) return rc;
            if (ranks && !fold) if (int rc = exchange_sum(x, bsum + (size_t)l * K * 8 * F, K * 8 * F, pdl_on(*desc), (cudaStream_t)stream)) return rc;
        }
    return gwtf_bwd_finish(desc, train, params, bnbuf, mom, bsum, gbuf, points, dparams, dpoints, B, N, n_total, stream);
}

int gwtf_bwd_all(const gwtf_stack_desc* desc, int32_t train, const float* params, const float* bnbuf,
                 const float* film, const float* points, const float* base, const float* logw, const float* ubuf,
                 const float* ybuf, const float* ld, const double* mom, const double* sum1, const float* nll,
                 const float* dnll,
                 double* bsum, float* gbuf, float* gs, float* dobuf, float* dparams, float* dfilm, float* dbase,
                 float* dlogw, float* dpoints, int32_t B, int32_t N, void* stream) {
    return bwd_all_impl(desc, train, params, bnbuf, film, points, base, logw, ubuf, ybuf, ld, mom, sum1, nll, dnll, bsum,
                        gbuf, gs, dobuf, dparams, dfilm, dbase, dlogw, dpoints, B, N, 0.0, stream);
}
int gwtf_bwd_all_ranks(const gwtf_stack_desc* desc, int32_t train, const float* params, const float* bnbuf,
                       const float* film, const float* points, const float* base, const float* logw, const float* ubuf,
                       const float* ybuf, const float* ld, const double* mom, const double* sum1, const float* nll,
                       const float* dnll, double* bsum, float* gbuf, float* gs, float* dobuf, float* dparams,
                       float* dfilm, float* dbase, float* dlogw, float* dpoints, int32_t B, int32_t N, double n_total,
                       void* stream) {
    if (!(n_total > 0.0)) return fail(-23, "n_total must be the number of points on all ranks");
    return bwd_all_impl(desc, train, params, bnbuf, film, points, base, logw, ubuf, ybuf, ld, mom, sum1, nll, dnll, bsum,
                        gbuf, gs, dobuf, dparams, dfilm, dbase, dlogw, dpoints, B, N, n_total, stream);
}

// ------------------------------------------------------------------------------------------ sampling
int gwtf_mixture_cdf(const float* logits, int32_t B, int32_t K, float* cdf, void* stream) {
    if (!logits || !cdf) return fail(-10, "null pointer argument");
    if (K < 1 || K > GWTF_MAX_COMPONENTS) return fail(-2, "n_components out of range");
    if (B <= 0) return 0;
    return launch_mixture_cdf(logits, B, K, cdf, (cudaStream_t)stream);
}

int gwtf_sample(const gwtf_stack_desc* desc, const float* params, const float* bnbuf, const float* film,
                const float* base, const float* cdf, int32_t B, int32_t N, uint64_t seed, uint32_t stream_id,
                const int32_t* idx_in, const float* eps_in, float* samples, int32_t* labels, float* z_out, void* stream) {
    if (int rc = check_desc(desc)) return rc;
    if (!params || !bnbuf || !film || !base || !cdf || !samples || !labels) return fail(-10, "null pointer argument");
    if (B <= 0 || N <= 0) return 0;
    SampleArgs a;
    a.d = *desc; a.params = params; a.bnbuf = bnbuf; a.film = film; a.base = base; a.cdf = cdf; a.B = B; a.N = N;
    a.seed_lo = (uint32_t)(seed & 0xFFFFFFFFull);
    a.stream_id = stream_id ^ (uint32_t)(seed >> 32);
    a.idx_in = idx_in; a.eps_in = eps_in; a.samples = samples; a.labels = labels; a.z_out = z_out;
    a.tiles_per_shape = 0; a.tile_points = 0;
    return launch_sample_fma(a, (cudaStream_t)stream);
}

static int64_t sample_npad(int K, int N) { return (int64_t)(N + 127) / 128 * 128 + 128 * (int64_t)K; }
static int64_t up4(int64_t n) { return (n + 3) / 4 * 4; }      // 16-byte aligned sections (in 4-byte words)

int64_t gwtf_sample_workspace_bytes(const gwtf_stack_desc* desc, int32_t B, int32_t N) {
    if (check_desc(desc) || B <= 0 || N <= 0) return 0;
    const int64_t K = desc->n_components, npad = sample_npad((int)K, N);
    // two ping-pong coordinate rows | slot (B,N) | counts, cursor (B,K) | seg (K,B,2) | seg_tiles (K,B+1)
    return 4 * (2 * B * 3 * npad + up4((int64_t)B * N) + 2 * up4(B * K) + up4(2 * K * B) + up4(K * (B + 1))) + 64;
}

int gwtf_sample_layers(const gwtf_stack_desc* desc, const float* params, const float* bnbuf, const float* film,
                       const float* base, const float* cdf, int32_t B, int32_t N, uint64_t seed, uint32_t stream_id,
                       const int32_t* idx_in, const float* eps_in, void* workspace, int64_t workspace_bytes,
                       float* samples, int32_t* labels, float* z_out, void* stream) {
    if (int rc = check_desc(desc)) return rc;
    if (!params || !bnbuf || !film || !base || !cdf || !workspace || !samples || !labels)
        return fail(-10, "null pointer argument");
    if (B <= 0 || N <= 0) return 0;
    if (fwd_engine(*desc) != kEngineTcFwd) return fail(-4, "gwtf_sample_layers needs the tcgen05 forward (use gwtf_sample)");
    if (workspace_bytes < gwtf_sample_workspace_bytes(desc, B, N)) return fail(-25, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int K = desc->n_components, L = desc->n_layers;
    const int npad = (int)sample_npad(K, N);
    const size_t row = (size_t)B * 3 * npad;
    float* xbuf = (float*)workspace;
    int32_t* slot = (int32_t*)(xbuf + 2 * row);
    int32_t* counts = slot + up4((int64_t)B * N);
    int32_t* cursor = counts + up4((int64_t)B * K);
    int32_t* seg = cursor + up4((int64_t)B * K);             // read as int2: needs 8-byte alignment
    int32_t* seg_tiles = seg + up4((int64_t)2 * K * B);
    GWTF_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)B * K, st));
    const uint32_t seed_lo = (uint32_t)(seed & 0xFFFFFFFFull), sid = stream_id ^ (uint32_t)(seed >> 32);
    SamplePlanArgs pa;
    pa.K = K; pa.B = B; pa.N = N; pa.cdf = cdf; pa.idx_in = idx_in; pa.counts = counts; pa.seed_lo = seed_lo; pa.stream_id = sid;
    if (int rc = launch_sample_count(pa, st)) return rc;
    if (int rc = launch_sample_plan(K, B, counts, seg, seg_tiles, cursor, st)) return rc;
    SampleScatterArgs sa;
    sa.K = K; sa.B = B; sa.N = N; sa.Npad = npad; sa.base = base; sa.cdf = cdf; sa.idx_in = idx_in; sa.eps_in = eps_in;
    sa.seed_lo = seed_lo; sa.stream_id = sid; sa.seg = seg; sa.cursor = cursor;
    sa.xbuf = xbuf; sa.slot = slot; sa.labels = labels; sa.z_out = z_out;
    if (int rc = launch_sample_scatter(sa, st)) return rc;
    const float* xin = xbuf;
    for (int l = 0; l < L; ++l) {               // direct order (flows.py:150-160)
        float* xout = xbuf + (size_t)((l + 1) & 1) * row;
        LayerArgs a;
        a.d = *desc; a.layer = l; a.train = 0; a.direct = 1; a.params = params; a.bnbuf = bnbuf; a.film = film;
        a.xin = xin; a.xin_shared = 1; a.xout = xout; a.ld = nullptr; a.ssum = nullptr; a.trio = nullptr; a.y1out = nullptr;
        a.mom_in = nullptr; a.mom_out = nullptr; a.sum1 = nullptr; a.B = B; a.N = npad; a.n_total = 1.0;
        a.tiles_per_shape = 0; a.seg = seg; a.seg_tiles = seg_tiles;
        if (int rc = launch_fwd_layer_tc(a, 1, st)) return rc;
        xin = xout;
    }
    return launch_sample_gather(B, N, npad, xin, slot, samples, st);
}

// ------------------------------------------------------------------------------------------ optimizer
// Fused AMSGrad / Adam step over one flat buffer (lib/networks/optimizers.py:42-74): first / second moment EMAs,
// running max of the second moment, bias corrections, and the reference's weight decay that is added to the
// UPDATE un-scaled by lr (:69-72).  HBM-bound: 5 reads + 4 writes of 4 bytes per parameter.

int gwtf_adam_step(float* p, const float* g, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq, int64_t n,
                   double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step, void* stream) {
    if (!p || !g || !exp_avg || !exp_avg_sq) return fail(-10, "null pointer argument");
    if (n <= 0) return 0;
    if (step < 1) return fail(-28, "step counts from 1");
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = sqrt(1.0 - pow(beta2, (double)step));
    long long blocks = (n + 256 * 4 - 1) / (256 * 4);
    const long long cap = 8LL * num_sms();
    if (blocks > cap) blocks = cap;
    k_adam<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, exp_avg, exp_avg_sq, max_exp_avg_sq, (long long)n,
                                                               (float)lr, (float)beta1, (float)beta2, (float)eps,
                                                               (float)weight_decay, (float)bc1, (float)bc2);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------ probes

int gwtf_mma_peak_tflops(int32_t iters, double* tflops, void* stream) {
    if (!tflops || iters <= 0) return fail(-1, "bad arguments to gwtf_mma_peak_tflops");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = num_sms() * 2;
    float best = 0.f;
    if (int rc = time_probe([&](float* out) { k_mma_probe<<<grid, 256, 0, st>>>(out, iters); }, grid, best, st)) return rc;
    // one m16n8k8 = 2*16*8*8 flops per warp instruction, 16 per iteration, 8 warps per CTA
    *tflops = 2.0 * 16 * 8 * 8 * 16.0 * (double)iters * 8.0 * (double)grid / ((double)best * 1e-3) * 1e-12;
    return 0;
}

int gwtf_fma_peak_tflops(int32_t iters, double* tflops, void* stream) {
    if (!tflops || iters <= 0) return fail(-1, "bad arguments to gwtf_fma_peak_tflops");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = num_sms() * 4;
    float best = 0.f;
    if (int rc = time_probe([&](float* out) { k_ffma_probe<<<grid, 256, 0, st>>>(out, iters, 1.0001f, 1e-4f); }, grid, best, st))
        return rc;
    *tflops = 2.0 * 32.0 * (double)iters * 256.0 * (double)grid / ((double)best * 1e-3) * 1e-12;
    return 0;
}

}  // extern "C"
