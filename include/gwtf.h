/*
 * gwtf.h -- C ABI of libgwtf.so: hand-written sm_100a CUDA kernels for the mixture-of-
 * conditional-RealNVP-flows log-likelihood (forward + backward) and its sampling pass.
 *
 * The reference (janisgp/go_with_the_flows) has no FFI for this path: the path IS the Python
 * modules in lib/networks/ (flows.py:95-117 CondRealNVPFlow3D.forward, flows.py:150-160 Triple,
 * decoders.py:61-79 LocalCondRNVPDecoder.forward, models.py:153-207 one_flow_decode,
 * flow_mixture.py:122-179 decode, losses.py:88-137 FlowMixtureNLL.forward) executed as ~53k ATen
 * calls per forward.  Each entry point below names the reference lines whose arithmetic it
 * replaces.  The Python drop-in (go_with_the_flows_b200/networks) binds them with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success, <0 for an argument error, >0 = cudaError_t;
 *     gwtf_last_error_string() describes the last failure on the calling thread.
 *   - all pointers are DEVICE pointers to contiguous fp32 unless noted; `stream` is a
 *     cudaStream_t passed as void*; nothing is allocated, nothing synchronises the host.
 *   - points / coordinates are (B,3,N) channel-major exactly like the reference tensors.
 *
 * Layouts (K components, L coupling layers in DIRECT order l = 3*triple + nvp-1, F features)
 *   params  [K][L][rec_stride]   per layer, per net X in (mu, logvar), natural tensor layouts:
 *             W0 (F,k) | bn0.weight (F) | bn0.bias (F) | W1 (F,F) | W2 (w,F) | b2 (w)
 *             net X starts at X*(F*F+5F+w); k = #kept dims, w = 3-k  (gwtf_param_offsets)
 *   bnbuf   [K][L][8F]           per net: running_mean0 | running_var0 | running_mean1 | running_var1
 *   film    [B][K][L][2][2][F]   per net: s = eps+exp(cond_w(g)) | t = cond_b(g)   (flows.py:100-106)
 *   base    [B][2][3]            mu_base | logvar_base of the base Gaussian (models.py:169-193)
 *   logw    [B][K]               NORMALISED log mixture weights (losses.py:101-104)
 *   ubuf    [L][K][B][3][N]      slot l = OUTPUT of layer l in the NLL (inverse) pass = input of
 *                                layer l-1; slot 0 = base-space sample z  (kept for backward)
 *   ybuf    [L][K][2][F][B][N]   (optional) y1 = FiLM(BN1(sd1(.)) pre-activation of both nets kept by the apply
 *                                pass so that backward skips one F x F contraction per phase (296 B per
 *                                point/component/layer at F=37; pass NULL to recompute instead)
 *   mom     [L][K][16] double    sum x_d (3), sum x_d x_e (6, upper triangle) of layer l's input
 *   sum1    [L][K][2][2][F] dbl  sum h1, sum h1^2 per net/channel (BatchNorm statistics of sd1_bn)
 *   bstat   [L][K][2][4][F]      batch mean0 | biased var0 | mean1 | biased var1 actually used
 */
#ifndef GWTF_H_
#define GWTF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GWTF_MAX_LAYERS 96
#define GWTF_MAX_COMPONENTS 16
#define GWTF_MAX_FEATURES 64
#define GWTF_MOM_STRIDE 16

typedef struct gwtf_stack_desc {
    int32_t n_components;                   /* K */
    int32_t n_layers;                       /* L = 3 * n_flows */
    int32_t n_features;                     /* F */
    int32_t rec_stride;                     /* floats per (component, layer) record in `params` */
    uint8_t warp_mask[GWTF_MAX_LAYERS];     /* bit d set: dim d is warped by layer l (flows.py:18-23) */
} gwtf_stack_desc;

int gwtf_version(void);
/* Contraction engine of the per-layer kernels:
 *   3 = warp-level tensor-core fragments (mma.sync m16n8k8 tf32, 3xTF32) for the forward AND backward
 *       layer phases, feature widths <= 64 (what engines 1-2 fall back to for widths > 39);
 *   2 = tcgen05 tensor cores for the forward phases, persistent warp-specialised kernel (3xTF32, TMEM
 *       accumulators, feature widths <= 39) + the mma.sync backward (default);
 *   1 = tcgen05 forward with one tile per 128-thread CTA + the mma.sync backward;
 *   0 = FP32 FMA pipe everywhere;  -1 = re-read the GWTF_TC environment variable.
 * Returns the previous setting (NOT an error code). */
int gwtf_set_tensor_cores(int32_t enable);
/* Programmatic dependent launch of the layer kernels (their parameter staging overlaps the predecessor's
 * tail): 1 = on (default), 0 = ordinary launches, -1 = re-read GWTF_PDL.  Returns the previous setting. */
int gwtf_set_pdl(int32_t enable);
/* The engine currently selected (0..3, environment resolved). */
int gwtf_engine(void);
/* Size in floats of the optional kept-activation buffer (`ybuf` of gwtf_fwd_layer / gwtf_fwd_all /
 * gwtf_bwd_layer / gwtf_bwd_all) for B clouds of N points under the CURRENT engine; the layout is private
 * to the engine, so the forward and backward calls that share a buffer must run under the same setting.
 * The tensor-core engines keep the sd1 output (engine 3: before, engines 1-2: after sd1_bn + FiLM, flows.py:100-104)
 * in MMA-fragment order: L*K*2*B*ceil(N/256)*256*roundup8(F) floats. */
int64_t gwtf_keep_floats(const gwtf_stack_desc* desc, int32_t B, int32_t N);
const char* gwtf_last_error_string(void);

/* Record geometry shared with the Python packer.  offsets[0..5] = W0,bn0.weight,bn0.bias,W1,W2,b2
 * relative to the start of a net; *net_stride = F*F+5F+w. */
int gwtf_rec_stride(int32_t n_features);
int gwtf_param_offsets(int32_t n_features, int32_t n_warp, int32_t* offsets, int32_t* net_stride);

/* Device-side FP32 FMA throughput probe used by bench.py for the roofline denominator. */
int gwtf_fma_peak_tflops(int32_t iters, double* tflops, void* stream);
/* Dense TF32 throughput of the warp-level tensor-core path (mma.sync.m16n8k8, register fragments) the
 * backward kernels run on, measured by an issue-rate probe kernel; the 3xTF32 split executes three such
 * MMAs per fp32-grade product. */
int gwtf_mma_peak_tflops(int32_t iters, double* tflops, void* stream);

/* ---- fused eval-mode forward NLL  (flow_mixture.py:163-166 + losses.py:88-137, BN running stats)
 * One launch walks all K stacks for every point, keeps xyz + log-det in registers, streams the
 * layer records through shared memory with TMA bulk copies and does the K-way log-sum-exp
 * in-kernel.  Outputs: nll (B,N); optional logp (B,N,K), z (K,B,3,N), ssum (K,B,3,N)
 * (= sum over layers of logvar, per dim: what `sum(p_prior_logvars[1:])` holds). */
int gwtf_nll_fwd_eval(const gwtf_stack_desc* desc, const float* params, const float* bnbuf,
                      const float* film, const float* points, const float* base, const float* logw,
                      int32_t B, int32_t N, float* nll, float* logp, float* z, float* ssum,
                      void* stream);

/* ---- phased forward (train-mode batch statistics, or eval-mode when gradients are needed)
 * gwtf_fwd_moments : input moments of the first processed layer (l = L-1) from the data points.
 * gwtf_fwd_layer   : phase 0 = statistics of h1 (sd1_bn batch stats), phase 1 = apply the layer,
 *                    write ubuf[l], add the layer's logvar to ld (K,B,N) [and ssum (K,B,3,N) if
 *                    given], accumulate the next layer's input moments.  With train=0 only
 *                    phase 1 is needed.  n_total = number of points the statistics are over
 *                    (B*N summed over ranks) -- the caller all-reduces mom/sum1 between phases.
 * gwtf_fwd_all     : single-process driver: moments + all layers + bstat + nll. */
/* Eval-mode NLL (same result as gwtf_nll_fwd_eval) through the per-layer tensor-core kernels of the current
 * engine: scratch = 2*K*B*3*N floats, ld = K*B*N floats (both overwritten). */
int gwtf_nll_fwd_eval_layers(const gwtf_stack_desc* desc, const float* params, const float* bnbuf,
                             const float* film, const float* points, const float* base, const float* logw,
                             float* scratch, float* ld, int32_t B, int32_t N, float* nll, float* logp, void* stream);
int gwtf_fwd_moments(const gwtf_stack_desc* desc, const float* points, int32_t B, int32_t N,
                     double* mom, void* stream);
int gwtf_fwd_layer(const gwtf_stack_desc* desc, int32_t layer, int32_t phase, int32_t train,
                   const float* params, const float* bnbuf, const float* film, const float* points,
                   float* ubuf, float* ld, float* ssum, float* ybuf, double* mom, double* sum1,
                   int32_t B, int32_t N, double n_total, void* stream);
/* explicit-pointer form of gwtf_fwd_layer: `xin` is (K,B,3,N), or the (B,3,N) data cloud when
 * xin_shared=1; direct=1 applies flows.py:113 instead of :115; `trio` (K,3,B,3,N), if given,
 * receives (p_out, mu, logvar) of the layer -- the per-module list API is built on this. */
int gwtf_fwd_layer_ex(const gwtf_stack_desc* desc, int32_t layer, int32_t phase, int32_t train, int32_t direct,
                      const float* params, const float* bnbuf, const float* film, const float* xin,
                      int32_t xin_shared, float* xout, float* ld, float* ssum, float* trio, float* y1out,
                      const double* mom_in, double* mom_out, double* sum1, int32_t B, int32_t N,
                      double n_total, void* stream);
int gwtf_fwd_bstat(const gwtf_stack_desc* desc, const float* params, const double* mom,
                   const double* sum1, double n_total, float* bstat, void* stream);
int gwtf_nll_from_state(const gwtf_stack_desc* desc, const float* ubuf, const float* ld,
                        const float* base, const float* logw, int32_t B, int32_t N,
                        float* nll, float* logp, void* stream);
int gwtf_fwd_all(const gwtf_stack_desc* desc, int32_t train, const float* params, const float* bnbuf,
                 const float* film, const float* points, const float* base, const float* logw,
                 float* ubuf, float* ld, float* ssum, float* ybuf, double* mom, double* sum1, float* bstat,
                 int32_t B, int32_t N, float* nll, float* logp, void* stream);

/* ---- backward (autograd of everything above)
 * gwtf_bwd_seed   : responsibilities, dL/dz, dL/dS, base and mixture-weight gradients from
 *                   dnll (B,N).  gbuf (K,B,3,N) <- dL/dz; gs (K,B,3,N) <- dL/dS;
 *                   dbase (B,2,3), dlogw (B,K) are accumulated (+=).
 * gwtf_bwd_layer  : phase 0 = recompute the layer, form d(o_mu,o_lv) into dobuf (K,B,6,N),
 *                   FiLM / sd2 gradients and the sd1_bn backward sums; phase 1 = sd1/sd0
 *                   gradients, bn0 sums, input gradient.  Layers are visited l = 0 .. L-1.
 *                   bsum (L,K,2,4,F) double holds per net: sum dn1 | sum dn1*n1 | sum dy0 | sum dy0*hhat0
 *                   (the caller all-reduces it between phases when ranks share batch statistics).
 * gwtf_bwd_finish : lazy bn0 correction of the last layer's input gradient -> dpoints (B,3,N)
 *                   (summed over components), and the closed-form sd0.weight gradients.
 * gwtf_bwd_all    : single-process driver.
 * dparams has the layout of params; dfilm the layout of film; both are accumulated (+=). */
int gwtf_bwd_seed(const gwtf_stack_desc* desc, const float* ubuf, const float* ld, const float* base,
                  const float* logw, const float* nll, const float* dnll, int32_t B, int32_t N,
                  float* gbuf, float* gs, float* dbase, float* dlogw, void* stream);
int gwtf_bwd_layer(const gwtf_stack_desc* desc, int32_t layer, int32_t phase, int32_t train,
                   const float* params, const float* bnbuf, const float* film, const float* points,
                   const float* ubuf, const float* ybuf, const double* mom, const double* sum1, double* bsum,
                   float* gbuf, const float* gs, float* dobuf, float* dparams, float* dfilm,
                   int32_t B, int32_t N, double n_total, void* stream);
int gwtf_bwd_finish(const gwtf_stack_desc* desc, int32_t train, const float* params,
                    const float* bnbuf, const double* mom, const double* bsum, const float* gbuf,
                    const float* points, float* dparams, float* dpoints, int32_t B, int32_t N,
                    double n_total, void* stream);
int gwtf_bwd_all(const gwtf_stack_desc* desc, int32_t train, const float* params, const float* bnbuf,
                 const float* film, const float* points, const float* base, const float* logw,
                 const float* ubuf, const float* ybuf, const float* ld, const double* mom, const double* sum1,
                 const float* nll, const float* dnll, double* bsum, float* gbuf, float* gs,
                 float* dobuf, float* dparams, float* dfilm, float* dbase, float* dlogw,
                 float* dpoints, int32_t B, int32_t N, void* stream);

/* ---- several ranks (one process per GPU) sharing batch statistics: SyncBatchNorm semantics of the
 * reference's DistributedDataParallel training (train_ae.py:77-78, torch.nn.SyncBatchNorm.convert_sync_batchnorm)
 * The statistic sums of every layer phase are exchanged through NVLink peer memory by a push / flag /
 * add kernel (csrc/gwtf_exchange.cuh) instead of a collective call per phase.
 * gwtf_exchange_attach : recv[r] / flags[r] = device pointers, valid on THIS device, to rank r's receive
 *                        buffer (2*world*slot_doubles doubles) and flag array (world uint64, zeroed before the
 *                        first exchange and never written by the host afterwards); symmetric / IPC memory.
 *                        world = 1 detaches.  All ranks must issue the same sequence of exchanges.
 * gwtf_exchange_sum    : in-place sum over ranks of n <= slot_doubles doubles (stream ordered).
 * gwtf_fwd_all_ranks / gwtf_bwd_all_ranks : the single-call drivers with n_total = points on all ranks. */
int gwtf_exchange_attach(int32_t rank, int32_t world, void* const* recv, void* const* flags, int32_t slot_doubles);
int gwtf_exchange_world(void);
int gwtf_exchange_sum(double* data, int32_t n, void* stream);
int gwtf_fwd_all_ranks(const gwtf_stack_desc* desc, int32_t train, const float* params, const float* bnbuf,
                       const float* film, const float* points, const float* base, const float* logw,
                       float* ubuf, float* ld, float* ssum, float* ybuf, double* mom, double* sum1, float* bstat,
                       int32_t B, int32_t N, float* nll, float* logp, double n_total, void* stream);
int gwtf_bwd_all_ranks(const gwtf_stack_desc* desc, int32_t train, const float* params, const float* bnbuf,
                       const float* film, const float* points, const float* base, const float* logw,
                       const float* ubuf, const float* ybuf, const float* ld, const double* mom, const double* sum1,
                       const float* nll, const float* dnll, double* bsum, float* gbuf, float* gs,
                       float* dobuf, float* dparams, float* dfilm, float* dbase, float* dlogw,
                       float* dpoints, int32_t B, int32_t N, double n_total, void* stream);

/* ---- sampling (flow_mixture.py:141-177 + models.py:199-203, eval-mode BN, lifted to a batch)
 * Per point: Philox4x32-10 counter (n, b, call, 0), key (seed lo32, stream ^ seed hi32) -> u -> component via
 * cdf (B,K) with np.searchsorted(side='right'); Box-Muller noise -> z = mu_base +
 * exp(lv_base/2)*eps; the component's DIRECT stack; labels = component+1 (int32).
 * Optional: idx_in (B,N) int32 / eps_in (B,3,N) override the in-kernel draws (parity tests);
 * z_out (B,3,N) returns the base-space sample that was pushed through the flow. */
int gwtf_sample(const gwtf_stack_desc* desc, const float* params, const float* bnbuf,
                const float* film, const float* base, const float* cdf, int32_t B, int32_t N,
                uint64_t seed, uint32_t stream_id, const int32_t* idx_in, const float* eps_in,
                float* samples, int32_t* labels, float* z_out, void* stream);
/* The same sampling pass through the per-layer tensor-core kernels of the current engine: points regrouped by the
 * component they drew.  gwtf_sample_plan counts the draws (counts (B,K) int32, *nmax = largest count; device
 * memory); the caller reads *nmax back, rounds it up to a multiple of 128 (nmax_pad) and provides
 * scratch = 2*K*B*3*nmax_pad floats and slot = B*N int32.  Results equal gwtf_sample's (same Philox streams). */
int gwtf_sample_plan(const gwtf_stack_desc* desc, const float* cdf, int32_t B, int32_t N, uint64_t seed,
                     uint32_t stream_id, const int32_t* idx_in, int32_t* counts, int32_t* nmax, void* stream);
int gwtf_sample_layers(const gwtf_stack_desc* desc, const float* params, const float* bnbuf, const float* film,
                       const float* base, const float* cdf, int32_t B, int32_t N, int32_t nmax_pad, uint64_t seed,
                       uint32_t stream_id, const int32_t* idx_in, const float* eps_in, float* scratch, int32_t* slot,
                       float* samples, int32_t* labels, float* z_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GWTF_H_ */
