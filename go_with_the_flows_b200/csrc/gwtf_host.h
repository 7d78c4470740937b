// Host-side plumbing shared by the translation units of libgwtf.so: error reporting, launch helpers,
// engine resolution and the launcher prototypes each kernel family exports.  The library keeps no mutable
// process-global state: engine, launch flags, precision tier, the rank exchange and the non-finite counter all
// ride in the caller-owned gwtf_stack_desc (include/gwtf.h).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#include "gwtf_common.cuh"

namespace gwtf {

struct EvalArgs;
struct LayerArgs;
struct BwdArgs;
struct SampleArgs;
struct SamplePlanArgs;
struct SampleScatterArgs;
struct FinishArgs;
struct AdamArgs;

// ---- error reporting (thread-local message, gwtf_last_error_string)
char* err_buf();
inline int fail(int code, const char* fmt, const char* what = "") {
    snprintf(err_buf(), 512, fmt, what);
    return code;
}
inline int cuda_fail(cudaError_t e, const char* where) {
    snprintf(err_buf(), 512, "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
}
#define GWTF_CUDA(x)                                              \
    do {                                                          \
        cudaError_t e__ = (x);                                    \
        if (e__ != cudaSuccess) return ::gwtf::cuda_fail(e__, #x); \
    } while (0)

// ---- engines (gwtf_stack_desc.engine)
//   0 = FP32 FMA pipe everywhere
//   2 = tcgen05 persistent forward (3xTF32) + mma.sync backward
//   3 = mma.sync forward and backward
//   4 = tcgen05 forward and backward (default)
constexpr int kEngineFma = 0, kEngineTcFwd = 2, kEngineMma = 3, kEngineTc = 4, kEngineDefault = 4;
inline int engine_of(const gwtf_stack_desc& d) {
    const int e = d.engine;
    return (e == kEngineFma || e == kEngineTcFwd || e == kEngineMma || e == kEngineTc) ? e : kEngineDefault;
}
// forward engine for feature width F: the tcgen05 kernels need F + 1 <= 40, wider stacks take the mma.sync path
inline int fwd_engine(const gwtf_stack_desc& d) {
    const int e = engine_of(d);
    if (e == kEngineFma || e == kEngineMma) return e;
    return d.n_features <= 39 ? kEngineTcFwd : kEngineMma;
}
// backward engine: tcgen05 (engine 4, F <= 39), mma.sync fragments (any F <= 64), or FMA
inline int bwd_engine(const gwtf_stack_desc& d) {
    const int e = engine_of(d);
    if (e == kEngineFma) return kEngineFma;
    if (e == kEngineTc && d.n_features <= 39) return kEngineTc;
    return kEngineMma;
}
inline bool pdl_on(const gwtf_stack_desc& d) { return (d.flags & GWTF_FLAG_NO_PDL) == 0; }

inline int num_sms() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);   // cached by the runtime; cheap
    return sms;
}

template <typename KernelT>
int blocks_per_sm(KernelT kernel, size_t smem, int threads = kThreads) {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem);
    return n < 1 ? 1 : n;
}

template <typename KernelT>
cudaError_t allow_smem(KernelT kernel, size_t smem) {
    // ask for the full shared-memory carveout: the default preference sizes L1 vs shared from a
    // heuristic and the occupancy query then reports 1 CTA/SM for 60-80 KB blocks
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

// launch with the programmatic-dependent-launch attribute (kernels that call pdl_wait() before touching
// anything their predecessor wrote); `pdl` false launches them the ordinary way
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                       Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

inline int padded_features(int F) {
    const int opts[] = {8, 16, 24, 32, 36, 40, 48, 64};
    for (int o : opts) if (F <= o) return o;
    return -1;
}

// points per thread of the FMA kernels for a padded width (register budget: P*FP accumulators)
template <int FP> struct PointsPerThread { static constexpr int fwd = FP <= 40 ? 4 : 2; static constexpr int bwd = FP <= 40 ? 2 : 1; };

#define GWTF_DISPATCH_FP(F, CALL)                     \
    switch (::gwtf::padded_features(F)) {             \
        case 8:  { constexpr int FP = 8;  CALL; } break;  \
        case 16: { constexpr int FP = 16; CALL; } break;  \
        case 24: { constexpr int FP = 24; CALL; } break;  \
        case 32: { constexpr int FP = 32; CALL; } break;  \
        case 36: { constexpr int FP = 36; CALL; } break;  \
        case 40: { constexpr int FP = 40; CALL; } break;  \
        case 48: { constexpr int FP = 48; CALL; } break;  \
        case 64: { constexpr int FP = 64; CALL; } break;  \
        default: return ::gwtf::fail(-4, "unsupported feature width"); \
    }

// tcgen05 kernels: K extent (features + the constant-one channel, padded to 8) and N extent (padded to 16)
#define GWTF_DISPATCH_TC(F, CALL)                                         \
    switch ((F + 8) / 8) {                                                \
        case 1: { constexpr int FPK = 8,  FPN = 16; CALL; } break;        \
        case 2: { constexpr int FPK = 16, FPN = 16; CALL; } break;        \
        case 3: { constexpr int FPK = 24, FPN = 32; CALL; } break;        \
        case 4: { constexpr int FPK = 32, FPN = 32; CALL; } break;        \
        case 5: { constexpr int FPK = 40, FPN = 48; CALL; } break;        \
        default: return ::gwtf::fail(-4, "unsupported feature width for the tensor-core path"); \
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
        default: return ::gwtf::fail(-4, "unsupported feature width"); \
    }

// ---- launchers, one translation unit per kernel family (gwtf_l_*.cu)
int launch_eval_fma(const EvalArgs& a, cudaStream_t st);                         // gwtf_l_fwd_fma.cu
int launch_fwd_layer_fma(const LayerArgs& a, int phase, cudaStream_t st);
int launch_moments(const float* points, int B, int N, int K, double* mom, cudaStream_t st);
int launch_bstat(const gwtf_stack_desc& d, const float* params, const double* mom, const double* sum1, double n_total,
                 float* bstat, cudaStream_t st);
int launch_nll_from_state(const gwtf_stack_desc& d, int B, int N, const float* ubuf, const float* ld, const float* base,
                          const float* logw, float* nll, float* logp, cudaStream_t st);
int launch_fwd_layer_tc(const LayerArgs& a, int phase, cudaStream_t st);          // gwtf_l_fwd_tc.cu
int launch_fwd_layer_mma(const LayerArgs& a, int phase, cudaStream_t st);         // gwtf_l_fwd_mma.cu
int launch_bwd_seed(int K, int B, int N, const float* ubuf, const float* ld, const float* base, const float* logw,
                    const float* nll, const float* dnll, float* gbuf, float* gs, float* dbase, float* dlogw,
                    cudaStream_t st);                                             // gwtf_l_bwd_fma.cu
int launch_bwd_layer_fma(const BwdArgs& a, int phase, cudaStream_t st);
int launch_bwd_finish(const FinishArgs& fa, const gwtf_stack_desc& d, int train, const float* params, const double* mom,
                      const double* bsum, const float* gbuf, const float* points, float* dpoints, int B, int N,
                      double n_total, cudaStream_t st);
int launch_bwd_layer_mma(const BwdArgs& a, int phase, cudaStream_t st);           // gwtf_l_bwd_mma_{d,e}.cu
int launch_bwd_layer_d_mma(const BwdArgs& a, cudaStream_t st);
int launch_bwd_layer_e_mma(const BwdArgs& a, cudaStream_t st);
int launch_bwd_layer_tc(const BwdArgs& a, int phase, cudaStream_t st);            // gwtf_l_bwd_tc.cu
int launch_sample_fma(const SampleArgs& a, cudaStream_t st);                      // gwtf_l_sample.cu
int launch_mixture_cdf(const float* logits, int B, int K, float* cdf, cudaStream_t st);
int launch_sample_count(const SamplePlanArgs& a, cudaStream_t st);
int launch_sample_plan(int K, int B, const int32_t* counts, int32_t* seg, int32_t* seg_tiles, int32_t* cursor,
                       cudaStream_t st);
int launch_sample_scatter(const SampleScatterArgs& a, cudaStream_t st);
int launch_sample_gather(int B, int N, int Npad, const float* xin, const int32_t* slot, float* samples, cudaStream_t st);
size_t keep_layer_floats(const gwtf_stack_desc& d, int B, int N);

}  // namespace gwtf
