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

/* Execution options ride in the descriptor the caller owns: the library keeps NO mutable process-global
 * state (two stacks, or two streams, in one process never share anything) and is re-entrant per stream. */
#define GWTF_ENGINE_DEFAULT (-1)  /* = GWTF_ENGINE_TC */
#define GWTF_ENGINE_FMA 0         /* FP32 FMA pipe everywhere (reference engine for cross-checks) */
#define GWTF_ENGINE_TC_FWD 2      /* tcgen05 persistent forward (3xTF32) + mma.sync register-fragment backward */
#define GWTF_ENGINE_MMA 3         /* mma.sync m16n8k8 tf32 fragments, forward and backward (feature widths <= 64) */
#define GWTF_ENGINE_TC 4          /* tcgen05 / TMEM forward AND backward (feature widths <= 39; wider -> MMA) */
#define GWTF_FLAG_NO_PDL 1        /* launch the layer kernels without programmatic dependent launch */
#define GWTF_PRECISION_3XTF32 0   /* fp32-grade: A*B ~ Ahi*Bhi + Alo*Bhi + Ahi*Blo */
#define GWTF_PRECISION_TF32 1     /* single-pass TF32: eval-mode NLL without gradients and sampling only
                                   * (per-point log-likelihood within 1e-5 of fp32, see tests/test_gpu_precision.py) */

typedef struct gwtf_exchange gwtf_exchange;   /* opaque: peer-memory statistic exchange of one rank */

typedef struct gwtf_stack_desc {
    int32_t n_components;                   /* K */
    int32_t n_layers;                       /* L = 3 * n_flows */
    int32_t n_features;                     /* F */
    int32_t rec_stride;                     /* floats per (component, layer) record in `params` */
    uint8_t warp_mask[GWTF_MAX_LAYERS];     /* bit d set: dim d is warped by layer l (flows.py:18-23) */
    int32_t engine;                         /* GWTF_ENGINE_* */
    int32_t flags;                          /* GWTF_FLAG_* */
    int32_t eval_precision;                 /* GWTF_PRECISION_* of the no-grad eval NLL / sampling passes */
    int32_t reserved;
    gwtf_exchange* exchange;                /* rank exchange used by gwtf_*_all_ranks, or NULL */
    int32_t* nonfinite;                     /* DEVICE counter, += number of non-finite per-point NLLs written by
                                             * gwtf_nll_from_state / gwtf_fwd_all / gwtf_nll_fwd_eval* (the reference
                                             * stops without an update on a NaN loss, training.py:43-46), or NULL */
} gwtf_stack_desc;

int gwtf_version(void);
/* The engine a descriptor resolves to for the forward (which=0) / backward (which=1) layer kernels. */
int gwtf_resolved_engine(const gwtf_stack_desc* desc, int32_t which);
/* Size in floats of the optional kept-activation buffer (`ybuf` of gwtf_fwd_layer / gwtf_fwd_all /
 * gwtf_bwd_layer / gwtf_bwd_all) for B clouds of N points under the descriptor's engine; the layout is private
 * to the engine.  0 for the tcgen05 backward, which always recomputes.
 * mma.sync backward: L*K*2*B*ceil(N/256)*256*roundup8(F) floats in MMA-fragment order. */
int64_t gwtf_keep_floats(const gwtf_stack_desc* desc, int32_t B, int32_t N);
/* Workspace queries of the single-call drivers (bytes; the caller allocates, the library never does). */
int64_t gwtf_eval_layers_workspace_bytes(const gwtf_stack_desc* desc, int32_t B, int32_t N);
int64_t gwtf_sample_workspace_bytes(const gwtf_stack_desc* desc, int32_t B, int32_t N);
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
/* Eval-mode NLL (same result as gwtf_nll_fwd_eval) through the per-layer tensor-core kernels of the descriptor's
 * engine; workspace of gwtf_eval_layers_workspace_bytes (overwritten).  desc->eval_precision selects fp32-grade
 * 3xTF32 or single-pass TF32 contractions. */
int gwtf_nll_fwd_eval_layers(const gwtf_stack_desc* desc, const float* params, const float* bnbuf,
                             const float* film, const float* points, const float* base, const float* logw,
                             void* workspace, int64_t workspace_bytes, int32_t B, int32_t N, float* nll, float* logp,
                             void* stream);
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
 * add step (csrc/gwtf_exchange.cuh) instead of a collective call per phase: the tcgen05 layer kernels run
 * it in their own tail (the last CTA of the producing launch), the other engines in a one-CTA kernel.
 * gwtf_exchange_create : recv[r] / flags[r] = device pointers, valid on THIS device, to rank r's receive
 *                        buffer (4*world*slot_doubles doubles: [2][world][slot] 16-byte cells, each double with
 *                        two copies of the exchange's sequence number -- no separate flag, no system fence on the
 *                        critical path; GWTF_EXCHANGE_LL=0 selects the data / fence / flag protocol on the first
 *                        half of the same buffer) and flag array (32 uint64 -- entries [0, world)
 *                        are the flags, entry 31 of the rank's own array is the library's last-CTA ticket --
 *                        zeroed before the first exchange and never written by the host afterwards);
 *                        symmetric / IPC memory.
 *                        timeout_s: seconds a rank waits for its peers before the exchange kernel gives up with
 *                        a (sticky) launch failure instead of hanging (<= 0: 600 s, NCCL-like).  All ranks must issue the same
 *                        sequence of exchanges; the handle owns the sequence counter, so one handle serves
 *                        one stream of work at a time (create one per concurrent stream).
 * gwtf_exchange_sum    : in-place sum over ranks of n <= slot_doubles doubles (stream ordered).
 * gwtf_exchange_resync : after a failed call on one rank: every rank calls it (collectively, after a barrier
 *                        of the caller's) to realign the sequence counters to `seq`.
 * gwtf_fwd_all_ranks / gwtf_bwd_all_ranks : the single-call drivers with n_total = points on all ranks;
 *                        they use desc->exchange. */
int gwtf_exchange_create(int32_t rank, int32_t world, void* const* recv, void* const* flags, int32_t slot_doubles,
                         double timeout_s, gwtf_exchange** out);
int gwtf_exchange_destroy(gwtf_exchange* x);
int gwtf_exchange_world(const gwtf_exchange* x);
uint64_t gwtf_exchange_seq(const gwtf_exchange* x);
int gwtf_exchange_resync(gwtf_exchange* x, uint64_t seq);
int gwtf_exchange_sum(gwtf_exchange* x, double* data, int32_t n, void* stream);
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
/* cdf (B,K) <- inclusive CDF of softmax(logits (B,K)) the way np.random.choice builds it (flow_mixture.py:149-153):
 * fp32 probabilities (e = fp32(exp(fp64(logit)))), float64 cumsum, normalised, last entry pinned to 1.  On the
 * device, so the sampling call never reads anything back to the host. */
int gwtf_mixture_cdf(const float* logits, int32_t B, int32_t K, float* cdf, void* stream);
/* The same sampling pass through the per-layer tcgen05 kernels: the points of every shape are regrouped by the
 * component they drew (segments inside the shape's row, sized from upper bounds -- no host read-back), pushed
 * through the L direct layers, gathered back.  Results equal gwtf_sample's (same Philox streams, bit-exact labels).
 * workspace: gwtf_sample_workspace_bytes.  Needs the tcgen05 forward (feature widths <= 39). */
int gwtf_sample_layers(const gwtf_stack_desc* desc, const float* params, const float* bnbuf, const float* film,
                       const float* base, const float* cdf, int32_t B, int32_t N, uint64_t seed, uint32_t stream_id,
                       const int32_t* idx_in, const float* eps_in, void* workspace, int64_t workspace_bytes,
                       float* samples, int32_t* labels, float* z_out, void* stream);

/* ---- optimizer (lib/networks/optimizers.py:42-74): one fused AMSGrad / Adam step over a flat fp32 buffer.
 * exp_avg, exp_avg_sq (and max_exp_avg_sq, NULL = no AMSGrad) are updated in place; `step` counts from 1 (bias
 * corrections 1-beta1^step, sqrt(1-beta2^step)); weight_decay is added to the update un-scaled by lr (:69-72). */
int gwtf_adam_step(float* p, const float* g, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq, int64_t n,
                   double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step, void* stream);

/* ---- diagnostics (host only, no GPU needed): replay of the tile schedule of the persistent layer kernels -- the
 * same code the kernels run.  For a launch of `total_tiles` 128-point tiles (`tiles_per_shape` per shape) on
 * `grid_x` CTAs per component with `slots` tiles in flight per CTA (4: forward and backward phase 0, 2: backward
 * phase 1), writes for every tile its CTA, tile slot and sequence number within the CTA; returns the number of
 * tiles assigned (== total_tiles) or a negative error.  In the two-slot kernel consecutive sequence numbers of a
 * CTA must belong to different slots (they take turns on one operand buffer); tests/test_dropin_cpu.py checks it. */
int gwtf_debug_tile_schedule(int32_t total_tiles, int32_t tiles_per_shape, int32_t grid_x, int32_t slots,
                             int32_t* tile_cta, int32_t* tile_slot, int32_t* tile_seq);

#ifdef __cplusplus
}
#endif
#endif /* GWTF_H_ */
