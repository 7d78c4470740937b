// C ABI of libgwtf.so (see include/gwtf.h).  Host-side argument checks, template dispatch on
// the padded feature width, launches on the caller's stream.  No allocation, no host sync.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#include "gwtf_common.cuh"
#include "gwtf_fwd.cuh"
#include "gwtf_tc_fwd.cuh"
#include "gwtf_tc_persist.cuh"
#include "gwtf_bwd.cuh"
#include "gwtf_bwd_mma.cuh"
#include "gwtf_fwd_mma.cuh"
#include "gwtf_sample.cuh"
#include "gwtf_exchange.cuh"

using namespace gwtf;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* what = "") {
    snprintf(g_err, sizeof(g_err), fmt, what);
    return code;
}
int cuda_fail(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
}
#define GWTF_CUDA(x)                                         \
    do {                                                     \
        cudaError_t e__ = (x);                               \
        if (e__ != cudaSuccess) return cuda_fail(e__, #x);   \
    } while (0)

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
    return 0;
}

int g_use_tc = -1;    // -1: decide from the environment (GWTF_TC), 0 = FMA, 1 = tcgen05 (3 CTAs/SM),
                      // 2 = tcgen05 persistent warp-specialised forward (1 CTA/SM, 4 tiles in flight; default),
                      // 3 = warp-level mma.sync fragments for the forward too (the backward always uses them)

int tc_mode(int F) {
    if (g_use_tc < 0) {
        const char* e = getenv("GWTF_TC");
        g_use_tc = e ? (e[0] == '0' ? 0 : (e[0] == '1' ? 1 : (e[0] == '3' ? 3 : 2))) : 2;
    }
    return g_use_tc;
}
// forward engine for feature width F: the tcgen05 kernels need F + 1 <= 40, wider stacks take the mma.sync path
int fwd_engine(int F) {
    tc_mode(F);
    if (g_use_tc == 0 || g_use_tc == 3) return g_use_tc;
    return F <= 39 ? g_use_tc : 3;
}
bool use_tc(int F) { const int m = fwd_engine(F); return m == 1 || m == 2; }
bool use_mma_fwd(int F) { return fwd_engine(F) == 3; }
// backward contractions on mma.sync register fragments (any F <= 64) unless the FMA engine is selected
bool use_mma_bwd() { tc_mode(1); return g_use_tc != 0; }

// launch with the programmatic-dependent-launch attribute (kernels that call pdl_wait() before touching
// anything their predecessor wrote); GWTF_PDL=0 launches them the ordinary way
int g_pdl = -1;        // programmatic dependent launch: -1 = from the environment (GWTF_PDL, default on)
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    if (g_pdl < 0) g_pdl = (getenv("GWTF_PDL") && getenv("GWTF_PDL")[0] == '0') ? 0 : 1;
    const bool on = g_pdl != 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = on ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// peer-memory statistic exchange of this process (gwtf_exchange_attach)
struct ExchangeCtx {
    int rank = 0, world = 1, slot = 0;
    unsigned long long seq = 0;
    double* recv[kMaxRanks] = {};
    unsigned long long* flags[kMaxRanks] = {};
} g_xchg;

int exchange_sum(double* data, int n, cudaStream_t st) {
    if (g_xchg.world <= 1) return 0;
    if (n > g_xchg.slot) return fail(-20, "exchange slot too small for this stack");
    ExchangeArgs a;
    a.rank = g_xchg.rank; a.world = g_xchg.world; a.n = n; a.slot = g_xchg.slot; a.seq = ++g_xchg.seq; a.data = data;
    for (int r = 0; r < kMaxRanks; ++r) { a.recv[r] = g_xchg.recv[r]; a.flags[r] = g_xchg.flags[r]; }
    GWTF_CUDA(launch_pdl(k_exchange_sum, dim3(1), dim3(256), 0, st, a));
    return 0;
}

int padded_features(int F) {
    const int opts[] = {8, 16, 24, 32, 36, 40, 48, 64};
    for (int o : opts) if (F <= o) return o;
    return -1;
}

int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return sms;
}

template <typename KernelT>
int blocks_per_sm(KernelT kernel, size_t smem) {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kThreads, smem);
    return n < 1 ? 1 : n;
}

template <typename KernelT>
cudaError_t allow_smem(KernelT kernel, size_t smem) {
    // ask for the full shared-memory carveout: the default preference sizes L1 vs shared from a
    // heuristic and the occupancy query then reports 1 CTA/SM for 60-80 KB blocks
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

// points per thread for a padded width (register budget: P*FP accumulators)
template <int FP> struct PointsPerThread { static constexpr int fwd = FP <= 40 ? 4 : 2; static constexpr int bwd = FP <= 40 ? 2 : 1; };

#define GWTF_DISPATCH_FP(F, CALL)                     \
    switch (padded_features(F)) {                     \
        case 8:  { constexpr int FP = 8;  CALL; } break;  \
        case 16: { constexpr int FP = 16; CALL; } break;  \
        case 24: { constexpr int FP = 24; CALL; } break;  \
        case 32: { constexpr int FP = 32; CALL; } break;  \
        case 36: { constexpr int FP = 36; CALL; } break;  \
        case 40: { constexpr int FP = 40; CALL; } break;  \
        case 48: { constexpr int FP = 48; CALL; } break;  \
        case 64: { constexpr int FP = 64; CALL; } break;  \
        default: return fail(-4, "unsupported feature width"); \
    }

// ------------------------------------------------------------------------------------------
template <int FP>
int launch_eval(const EvalArgs& a0, cudaStream_t st) {
    constexpr int P = PointsPerThread<FP>::fwd;
    EvalArgs a = a0;
    a.tiles_per_shape = (a.N + kThreads * P - 1) / (kThreads * P);
    const int F = a.d.n_features;
    const size_t smem = round_up((int)sizeof(EvalSmem<FP>), 16) + 2 * (size_t)round_up(raw_floats(F), 4) * 4;
    GWTF_CUDA(allow_smem(k_nll_eval<FP, P>, smem));
    k_nll_eval<FP, P><<<a.B * a.tiles_per_shape, kThreads, smem, st>>>(a);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

#define GWTF_DISPATCH_TC(F, CALL)                                         \
    switch ((F + 8) / 8) {                                                \
        case 1: { constexpr int FPK = 8,  FPN = 16; CALL; } break;        \
        case 2: { constexpr int FPK = 16, FPN = 16; CALL; } break;        \
        case 3: { constexpr int FPK = 24, FPN = 32; CALL; } break;        \
        case 4: { constexpr int FPK = 32, FPN = 32; CALL; } break;        \
        case 5: { constexpr int FPK = 40, FPN = 48; CALL; } break;        \
        default: return fail(-4, "unsupported feature width for the tensor-core path"); \
    }

template <typename KernelT>
int blocks_per_sm_n(KernelT kernel, int threads, size_t smem) {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem);
    return n < 1 ? 1 : n;
}

template <int FPK, int FPN, int PHASE>
int launch_fwd_layer_tc(const LayerArgs& a0, cudaStream_t st) {
    LayerArgs a = a0;
    a.tiles_per_shape = (a.N + kTcThreads - 1) / kTcThreads;
    const int F = a.d.n_features;
    const size_t smem = round_up((int)sizeof(TcFwdSmem<FPK, FPN>), 16) + (size_t)round_up(raw_floats(F), 4) * 4;
    auto kern = k_fwd_layer_tc<FPK, FPN, PHASE>;
    GWTF_CUDA(allow_smem(kern, smem));
    const int tiles = a.B * a.tiles_per_shape;
    const int K = a.d.n_components;
    // the occupancy API reports 1 CTA/SM for tcgen05 kernels; size the grid from the real limits
    // (shared memory, tensor-memory columns) -- the hardware co-schedules what fits
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    (void)blocks_per_sm_n(kern, kTcThreads, smem);
    if (getenv("GWTF_DEBUG")) {
        static int once = 0;
        if (!once++) fprintf(stderr, "[gwtf] tc fwd: occupancy %d CTAs/SM, smem %zu, sms %d, err=%s\n", per_sm, smem, num_sms(), cudaGetErrorString(cudaPeekAtLastError()));
    }
    if (per_sm > 512 / kTcCols) per_sm = 512 / kTcCols;          // tensor-memory columns per SM
    int gx = (num_sms() * per_sm + K - 1) / K;
    if (gx > tiles) gx = tiles;
    if (gx < 1) gx = 1;
    kern<<<dim3(gx, K), kTcThreads, smem, st>>>(a);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

template <int FPK, int FPN, int PHASE>
int launch_fwd_layer_tcp(const LayerArgs& a0, cudaStream_t st) {
    LayerArgs a = a0;
    a.tiles_per_shape = (a.N + 127) / 128;
    const int F = a.d.n_features, K = a.d.n_components;
    const size_t smem = round_up((int)sizeof(TcPersistSmem<FPK, FPN>), 16) + (size_t)round_up(raw_floats(F), 4) * 4;
    auto kern = k_fwd_layer_tcp<FPK, FPN, PHASE>;
    GWTF_CUDA(allow_smem(kern, smem));
    const int tiles = a.B * a.tiles_per_shape;
    int gx = num_sms() / K;                                  // one CTA per SM owns all 512 TMEM columns
    if (gx > (tiles + kSlots - 1) / kSlots) gx = (tiles + kSlots - 1) / kSlots;
    if (gx < 1) gx = 1;
    GWTF_CUDA(launch_pdl(kern, dim3(gx, K), dim3(kPersistThreads), smem, st, a));
    return 0;
}

template <int FP, int PHASE>
int launch_fwd_layer_mma(const LayerArgs& a, cudaStream_t st) {
    // statistics pass: 2 m-tiles per warp (shared B fragments, 2 MMA chains), full register file, 1 CTA/SM;
    // apply pass: 1 m-tile per warp, 128 registers, 2 CTAs/SM
    constexpr int MI = PHASE == 0 ? 2 : 1;
    const int F = a.d.n_features, K = a.d.n_components;
    const size_t smem = fwd_mma_smem<FP>(F);
    auto kern = k_fwd_layer_mma<FP, PHASE, MI>;
    GWTF_CUDA(allow_smem(kern, smem));
    const long long tiles = (long long)a.B * ((a.N + 128 * MI - 1) / (128 * MI));
    int gx = (MI == 1 ? 2 : 1) * num_sms() / K;              // contiguous tile ranges
    if (gx > tiles) gx = (int)tiles;
    kern<<<dim3(gx < 1 ? 1 : gx, K), kThreads, smem, st>>>(a, PHASE == 1 ? a.y1out : nullptr);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

#define GWTF_DISPATCH_FP8(F, CALL)                        \
    switch (((F) + 7) / 8 * 8) {                          \
        case 8:  { constexpr int FP = 8;  CALL; } break;  \
        case 16: { constexpr int FP = 16; CALL; } break;  \
        case 24: { constexpr int FP = 24; CALL; } break;  \
        case 32: { constexpr int FP = 32; CALL; } break;  \
        case 40: { constexpr int FP = 40; CALL; } break;  \
        case 48: { constexpr int FP = 48; CALL; } break;  \
        case 56: { constexpr int FP = 56; CALL; } break;  \
        case 64: { constexpr int FP = 64; CALL; } break;  \
        default: return fail(-4, "unsupported feature width"); \
    }

// floats per layer of the kept-activation buffer under the current engine
size_t keep_layer_floats(int F, int K, int B, int N) {
    return fwd_engine(F) != 0 ? (size_t)K * mma_keep_floats(F, B, N) : (size_t)K * 2 * F * B * N;
}

template <int FP, int PHASE>
int launch_fwd_layer(const LayerArgs& a0, cudaStream_t st) {
    constexpr int P = PointsPerThread<FP>::fwd;
    LayerArgs a = a0;
    a.tiles_per_shape = (a.N + kThreads * P - 1) / (kThreads * P);
    const int F = a.d.n_features;
    const size_t smem = round_up((int)sizeof(PhaseSmem<FP>), 16) + (size_t)round_up(raw_floats(F), 4) * 4;
    auto kern = k_fwd_layer<FP, P, PHASE>;
    GWTF_CUDA(allow_smem(kern, smem));
    const int tiles = a.B * a.tiles_per_shape;
    const int K = a.d.n_components;
    int gx = (num_sms() * blocks_per_sm(kern, smem) + K - 1) / K;
    if (gx > tiles) gx = tiles;
    if (gx < 1) gx = 1;
    kern<<<dim3(gx, K), kThreads, smem, st>>>(a);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

extern "C" {

int gwtf_version(void) { return 1; }

int gwtf_set_tensor_cores(int32_t enable) {
    const int prev = g_use_tc;
    g_use_tc = enable < 0 ? -1 : (enable > 3 ? 3 : enable);
    return prev;
}
const char* gwtf_last_error_string(void) { return g_err; }

int gwtf_set_pdl(int32_t enable) {
    const int prev = g_pdl;
    g_pdl = enable < 0 ? -1 : (enable ? 1 : 0);
    return prev;
}
int gwtf_engine(void) { tc_mode(1); return g_use_tc; }

int64_t gwtf_keep_floats(const gwtf_stack_desc* desc, int32_t B, int32_t N) {
    if (check_desc(desc) || B <= 0 || N <= 0) return 0;
    return (int64_t)desc->n_layers * (int64_t)keep_layer_floats(desc->n_features, desc->n_components, B, N);
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
    GWTF_DISPATCH_FP(desc->n_features, return launch_eval<FP>(a, (cudaStream_t)stream));
    return 0;
}

int gwtf_fwd_moments(const gwtf_stack_desc* desc, const float* points, int32_t B, int32_t N, double* mom, void* stream) {
    if (int rc = check_desc(desc)) return rc;
    if (!points || !mom) return fail(-10, "null pointer argument");
    if (B <= 0 || N <= 0) return 0;
    size_t total = (size_t)B * N;
    int grid = (int)((total + kThreads * 8 - 1) / (kThreads * 8));
    if (grid > 4 * num_sms()) grid = 4 * num_sms();
    k_moments<<<grid, kThreads, 0, (cudaStream_t)stream>>>(points, B, N, desc->n_components, mom);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

// NOTE: the C ABI takes explicit per-layer pointers so a multi-rank caller can all-reduce
// mom / sum1 between phases; xin_shared=1 when `xin` is the (B,3,N) data cloud.
int gwtf_fwd_layer_ex(const gwtf_stack_desc* desc, int32_t layer, int32_t phase, int32_t train, int32_t direct,
                      const float* params, const float* bnbuf, const float* film, const float* xin,
                      int32_t xin_shared, float* xout, float* ld, float* ssum, float* trio, float* y1out,
                      const double* mom_in, double* mom_out, double* sum1, int32_t B, int32_t N, double n_total,
                      void* stream) {
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
    if (use_mma_fwd(desc->n_features)) {
        if (phase == 0) { GWTF_DISPATCH_FP8(desc->n_features, return (launch_fwd_layer_mma<FP, 0>(a, (cudaStream_t)stream))); }
        else { GWTF_DISPATCH_FP8(desc->n_features, return (launch_fwd_layer_mma<FP, 1>(a, (cudaStream_t)stream))); }
        return 0;
    }
    if (fwd_engine(desc->n_features) == 2 && desc->n_components <= num_sms()) {
        if (phase == 0) { GWTF_DISPATCH_TC(desc->n_features, return (launch_fwd_layer_tcp<FPK, FPN, 0>(a, (cudaStream_t)stream))); }
        else { GWTF_DISPATCH_TC(desc->n_features, return (launch_fwd_layer_tcp<FPK, FPN, 1>(a, (cudaStream_t)stream))); }
        return 0;
    }
    if (use_tc(desc->n_features)) {
        if (phase == 0) { GWTF_DISPATCH_TC(desc->n_features, return (launch_fwd_layer_tc<FPK, FPN, 0>(a, (cudaStream_t)stream))); }
        else { GWTF_DISPATCH_TC(desc->n_features, return (launch_fwd_layer_tc<FPK, FPN, 1>(a, (cudaStream_t)stream))); }
        return 0;
    }
    if (phase == 0) { GWTF_DISPATCH_FP(desc->n_features, return (launch_fwd_layer<FP, 0>(a, (cudaStream_t)stream))); }
    else { GWTF_DISPATCH_FP(desc->n_features, return (launch_fwd_layer<FP, 1>(a, (cudaStream_t)stream))); }
    return 0;
}

int gwtf_fwd_layer(const gwtf_stack_desc* desc, int32_t layer, int32_t phase, int32_t train, const float* params,
                   const float* bnbuf, const float* film, const float* points, float* ubuf, float* ld, float* ssum,
                   float* ybuf, double* mom, double* sum1, int32_t B, int32_t N, double n_total, void* stream) {
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
    float* y1 = ybuf ? ybuf + (size_t)layer * keep_layer_floats(F, K, B, N) : nullptr;
    return gwtf_fwd_layer_ex(desc, layer, phase, train, 0, params, bnbuf, film, xin, first ? 1 : 0,
                             ubuf + (size_t)layer * slot, ld, ssum, nullptr, y1, mom_in, train ? mom_out : nullptr, s1,
                             B, N, n_total, stream);
}

int gwtf_fwd_bstat(const gwtf_stack_desc* desc, const float* params, const double* mom, const double* sum1,
                   double n_total, float* bstat, void* stream) {
    if (int rc = check_desc(desc)) return rc;
    if (!params || !mom || !sum1 || !bstat) return fail(-10, "null pointer argument");
    const int total = desc->n_layers * desc->n_components * 2 * desc->n_features;
    k_bstat<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*desc, params, mom, sum1, n_total, bstat);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

int gwtf_nll_from_state(const gwtf_stack_desc* desc, const float* ubuf, const float* ld, const float* base,
                        const float* logw, int32_t B, int32_t N, float* nll, float* logp, void* stream) {
    if (int rc = check_desc(desc)) return rc;
    if (!ubuf || !ld || !base || !logw || !nll) return fail(-10, "null pointer argument");
    if (B <= 0 || N <= 0) return 0;
    const size_t total = (size_t)B * N;
    k_nll_from_state<<<(unsigned)((total + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        desc->n_components, B, N, ubuf, ld, base, logw, nll, logp);
    GWTF_CUDA(cudaGetLastError());
    return 0;
}

// Eval-mode NLL through the per-layer kernels of the current engine (tensor cores): L launches that ping-pong
// between two (K,B,3,N) slots of `scratch`, then the mixture head.  Faster than the single-launch FMA kernel
// at every size measured on B200 (64 x 2048: 2.2 vs 3.3 ms; 4 x 2048: 0.5 vs 3.3 ms).
int gwtf_nll_fwd_eval_layers(const gwtf_stack_desc* desc, const float* params, const float* bnbuf, const float* film,
                             const float* points, const float* base, const float* logw, float* scratch, float* ld,
                             int32_t B, int32_t N, float* nll, float* logp, void* stream) {
    if (int rc = check_desc(desc)) return rc;
    if (!params || !bnbuf || !film || !points || !base || !logw || !scratch || !ld || !nll)
        return fail(-10, "null pointer argument");
    if (B <= 0 || N <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int L = desc->n_layers, K = desc->n_components;
    const size_t slot = (size_t)K * B * 3 * N;
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

int gwtf_exchange_attach(int32_t rank, int32_t world, void* const* recv, void* const* flags, int32_t slot_doubles) {
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) return fail(-21, "bad rank / world size");
    if (world > 1 && (!recv || !flags || slot_doubles <= 0)) return fail(-21, "exchange buffers missing");
    g_xchg = ExchangeCtx();
    g_xchg.rank = rank; g_xchg.world = world; g_xchg.slot = slot_doubles;
    for (int r = 0; r < world && world > 1; ++r) {
        if (!recv[r] || !flags[r]) return fail(-21, "null peer pointer");
        g_xchg.recv[r] = (double*)recv[r];
        g_xchg.flags[r] = (unsigned long long*)flags[r];
    }
    return 0;
}
int gwtf_exchange_world(void) { return g_xchg.world; }
int gwtf_exchange_sum(double* data, int32_t n, void* stream) {
    if (!data || n <= 0) return fail(-10, "null pointer argument");
    return exchange_sum(data, n, (cudaStream_t)stream);
}

// n_total > 0: statistics are over n_total points on all ranks; the sums are exchanged between phases
// through the attached peer-memory exchange
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
    if (ranks && g_xchg.world <= 1) return fail(-22, "gwtf_exchange_attach has not been called");
    const double n_total = n_total_in > 0.0 ? n_total_in : (double)B * (double)N;
    GWTF_CUDA(cudaMemsetAsync(ld, 0, sizeof(float) * (size_t)K * B * N, st));
    if (ssum) GWTF_CUDA(cudaMemsetAsync(ssum, 0, sizeof(float) * (size_t)K * B * 3 * N, st));
    if (train) {
        if (!mom || !sum1) return fail(-10, "train mode needs mom and sum1");
        GWTF_CUDA(cudaMemsetAsync(mom, 0, sizeof(double) * (size_t)L * K * GWTF_MOM_STRIDE, st));
        GWTF_CUDA(cudaMemsetAsync(sum1, 0, sizeof(double) * (size_t)L * K * 4 * F, st));
        if (int rc = gwtf_fwd_moments(desc, points, B, N, mom + (size_t)(L - 1) * K * GWTF_MOM_STRIDE, stream)) return rc;
    }
    for (int l = L - 1; l >= 0; --l) {
        if (train) {
            if (ranks) if (int rc = exchange_sum(mom + (size_t)l * K * GWTF_MOM_STRIDE, K * GWTF_MOM_STRIDE, st)) return rc;
            if (int rc = gwtf_fwd_layer(desc, l, 0, 1, params, bnbuf, film, points, ubuf, ld, ssum, ybuf, mom, sum1, B,
                                        N, n_total, stream)) return rc;
            if (ranks) if (int rc = exchange_sum(sum1 + (size_t)l * K * 4 * F, K * 4 * F, st)) return rc;
        }
        if (int rc = gwtf_fwd_layer(desc, l, 1, train, params, bnbuf, film, points, ubuf, ld, ssum, ybuf, mom, sum1, B,
                                    N, n_total, stream)) return rc;
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

#include "gwtf_api_bwd.inc"

}  // extern "C"
