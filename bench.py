#!/usr/bin/env python
"""Benchmark of the mixture-of-flows hot path (BASELINE.json metric: points/sec fwd+bwd mixture NLL and sampling).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W     # the reference's own modules on the host cores

Headline (`value`, `e2e`): a step = one train-mode (batch-statistics BatchNorm) forward + backward of the
K-component mixture-of-flows NLL over one synthetic batch: `Flow_Mixture_Model.decode` + `FlowMixtureNLL` +
`.backward()`, i.e. everything between the shape latent / point cloud and the gradients of every decoder
parameter, the latent, the base Gaussian and the mixture weights.  Workload = BASELINE.json configs[1] (C2):
config_generative_modeling_airplane.yaml model (K=4, 33 coupling layers, F=37, G=128), 64 clouds x 2048 points
per GPU (weak scaling), random init (seed 0), synthetic clouds N(0, 0.2^2), latents N(0, 0.5^2).

`extra` holds the other BASELINE configs, each run on EVERY rank (shapes sharded, no collective unless it is a
train step) with the aggregate over ranks and its own roofline:
  eval_nll      C1-shaped eval-mode NLL (fp32-grade and the single-pass TF32 tier)
  c3_step       config_autoencoding.yaml (F=33, G=512, freevar) train step
  small_step    C2 model at 4 clouds x 2048 per GPU: the launch-bound regime
  strong        C2 at the reference's semantics: 64 clouds in total, 64/N per GPU (train_ae.py:77-78)
  sampling_c4   config_SVR.yaml decoder: 256 latents x 2048 points and 64 x 2500
  sampling_c5   sweep 2k / 16k / 128k / 1M points x 256 latents, airplane decoder

Roofline accounting: algorithmic contraction FLOPs over CUDA-event time.
Tensor kernels are rated against MEASURED_PEAKS.json `bf16_tflops` / 2 (dense TF32 runs at half the bf16 rate) / 3
(fp32-grade 3xTF32 issues three MMAs per product) -- "of measured"; `frac_of_fma_peak` rates the same number
against the FP32 FMA pipe (live FFMA probe), which is what the >= 50 % target of the north star is quoted on.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def flops_per_point(F, K, L):
    """Algorithmic contraction FLOPs: per point, component, layer 2 nets x
    (F*k + F*F + F*w) MAC = 4(F^2+3F) FLOP forward; backward (dgrad + wgrad) = 2x forward."""
    fwd = K * L * 4 * (F * F + 3 * F)
    return fwd, 3 * fwd


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return {'bf16_tflops': float(d['bf16_tflops']), 'bf16_tflops_sustained': float(d.get('bf16_tflops_sustained', d['bf16_tflops'])),
                'hbm_gbs': float(d['hbm_gbs']), 'source': 'measured'}
    return {'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'hbm_gbs': 6650.0, 'source': 'fallback'}


def build_model(cfg_name, device):
    from go_with_the_flows_b200 import configs
    from go_with_the_flows_b200.networks.flow_mixture import Flow_Mixture_Model
    cfg = dict(configs.BY_NAME[cfg_name])
    torch.manual_seed(0)
    model = Flow_Mixture_Model(**cfg)
    return cfg, model.to(device)


def synthetic(B, N, G, seed_shift=0):
    gen = torch.Generator().manual_seed(1234 + seed_shift)
    p = 0.2 * torch.randn(B, 3, N, generator=gen)
    gen = torch.Generator().manual_seed(4321 + seed_shift)
    g = 0.5 * torch.randn(B, G, generator=gen)
    return p, g


class ClockSampler:
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.QUERY, '--format=csv,noheader,nounits',
                 '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arms: the reference's own modules (staged by build() into oracle/_ref, git-ignored) or the oracle port
# ---------------------------------------------------------------------------------------------
def _reference_modules():
    ref_root = os.path.join(ROOT, 'oracle', '_ref')
    if not os.path.isdir(os.path.join(ref_root, 'lib', 'networks')):
        return None
    sys.dont_write_bytecode = True
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    try:
        from lib.networks.flow_mixture import Flow_Mixture_Model as RefModel
        from lib.networks.losses import FlowMixtureNLL as RefNLL
        return RefModel, RefNLL
    except Exception:
        return None


def cpu_step_time(cfg_name, B, N, steps, warmup):
    """Train-mode forward + backward of the mixture NLL on the host cores, bounded sample of the workload.
    -> (seconds per step, threads, kind)."""
    torch.set_num_threads(os.cpu_count() or 1)
    from go_with_the_flows_b200 import configs
    cfg = dict(configs.BY_NAME[cfg_name])
    p, g = synthetic(B, N, cfg['g_latent_space_size'])
    ref = _reference_modules()
    if ref is not None:
        RefModel, RefNLL = ref
        torch.manual_seed(0)
        model = RefModel(**cfg)
        model.mode = 'training'
        model.train()
        loss_fn = RefNLL()
        g.requires_grad_(True)

        def step():
            model.zero_grad()
            out, logits = model.decode(p, g, N, False, False)
            loss = loss_fn(out, logits)
            loss.backward()
            return float(loss)
        kind = 'reference'
    else:
        from oracle import flow_oracle as fo
        _, model = build_model(cfg_name, 'cpu')
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        for k, v in sd.items():
            if v.is_floating_point() and k.startswith(('pc_decoder', 'p_prior', 'mixture_weights')) \
                    and 'running' not in k and not k.endswith('eps'):
                v.requires_grad_(True)
        g.requires_grad_(True)

        def step():
            for v in sd.values():
                v.grad = None
            out = fo.mixture_nll(p, g, sd, base_type=cfg['p_decoder_base_type'], weights_type=cfg['weights_type'],
                                 training=True, base_var=cfg['p_decoder_base_var'])
            out['pnll'].backward()
            return float(out['pnll'])
        kind = 'port'
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return sum(times) / len(times), torch.get_num_threads(), kind


def workload_name(args):
    if args.config == 'generative':
        return 'C2: config_generative_modeling_airplane model (K=4,L=33,F=37,G=128), train-mode fwd+bwd mixture NLL, ' \
               '%d clouds x %d points per GPU' % (args.batch, args.points)
    return '%s config, train-mode fwd+bwd mixture NLL, %d clouds x %d points per GPU' % (args.config, args.batch, args.points)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    B, N = 4, args.points
    steps = max(1, min(args.steps, 3))
    warmup = max(1, min(args.warmup, 1))
    sec, cores, kind = cpu_step_time(args.config, B, N, steps, warmup)
    value = B * N / sec
    sample = 'train-mode fwd+bwd mixture NLL, %d clouds x %d points per step (autograd holds ~1 GB/shape; B=4 is the ' \
             'CPU path\'s best per-point operating point), %d timed steps after %d warm-up' % (B, N, steps, warmup)
    line = {
        'impl': 'reference', 'metric': 'points/sec fwd+bwd mixture-flow NLL', 'value': value, 'unit': 'points/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args), 'cpu_sample_clouds': B},
        'cpu_baseline': {'value': value, 'unit': 'points/s', 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': 'points/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# device-side timing helpers
# ---------------------------------------------------------------------------------------------
class Env:
    def __init__(self, world, rank, dev):
        self.world, self.rank, self.dev = world, rank, dev
        self.flush_buf = torch.empty(256 * 1024 * 1024 // 4, device=dev)

    def barrier(self):
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def timed(self, fn, reps, warmup=1, flush=True):
        """Mean CUDA-event time of `fn` over `reps` calls (L2 flushed before each), max over ranks."""
        for _ in range(warmup):
            fn()
        self.barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for e0, e1 in evs:
            if flush:
                self.flush_buf.fill_(0.0)
            e0.record()
            fn()
            e1.record()
        self.barrier()
        return self.max_over_ranks(sum(e0.elapsed_time(e1) for e0, e1 in evs) / reps)


def make_step(model, world, N):
    from go_with_the_flows_b200.networks.losses import FlowMixtureNLL
    loss_fn = FlowMixtureNLL()
    stack = model.flow_stack()
    stack.prepare()
    params_all = list(model.parameters())
    owned = set()
    for name in stack.grad_masters:          # decoder tensors: averaged across ranks inside backward
        owned.update(id(q) for q in stack.masters[name].params)
    params_other = [q for q in params_all if id(q) not in owned]

    def step(p, g):
        for q in params_all:
            q.grad = None
        g = g.detach().requires_grad_(True)
        out_dec, logits = model.decode(p, g, N)
        pnll = loss_fn(out_dec, logits)
        pnll.backward()
        if world > 1:
            grads = [q.grad for q in params_other if q.grad is not None]
            flat = torch._utils._flatten_dense_tensors(grads)
            dist.all_reduce(flat)
            flat.div_(world)
            for gr, fl in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
                gr.copy_(fl)
        return pnll
    return step


def kernel_only_times(model, p, g, reps):
    """CUDA-event time of the step's own kernels with nothing in between: forward driver
    (moments + 2 phases x L layers + bstat + nll) and backward driver (seed + 2 phases x L + finish),
    plus each phase class timed over the L layers (ordinary launches)."""
    from go_with_the_flows_b200 import _native as nat
    from go_with_the_flows_b200.flowstack import _stream_ptr
    lib = nat.lib()
    stack = model.flow_stack()
    K, L, Fd = stack.K, stack.L, stack.F
    B, _, N = p.shape
    dev = p.device
    with torch.no_grad():
        film = stack.film(g, False, False)
        params = stack.pack_params().contiguous()
        bnbuf = stack.pack_bn()
        mu_b, lv_b = model.base_gaussian(g)
        base = torch.stack([mu_b, lv_b], 1).contiguous()
        logits = model.get_weights(g)
        logw = (logits - torch.logsumexp(logits, -1, keepdim=True)).contiguous()
    ubuf = torch.empty(L, K, B, 3, N, device=dev)
    ld = torch.zeros(K, B, N, device=dev)
    mom = torch.zeros(L, K, nat.MOM_STRIDE, device=dev, dtype=torch.float64)
    sum1 = torch.zeros(L, K, 2, 2, Fd, device=dev, dtype=torch.float64)
    bstat = torch.empty(L, K, 2, 4, Fd, device=dev)
    nll = torch.empty(B, N, device=dev)
    dnll = torch.full((B, N), 1.0 / B, device=dev)
    bsum = torch.zeros(L, K, 2, 4, Fd, device=dev, dtype=torch.float64)
    gbuf = torch.empty(K, B, 3, N, device=dev)
    gs = torch.empty(K, B, 3, N, device=dev)
    dobuf = torch.empty(K, B, 6, N, device=dev)
    dparams = torch.zeros_like(params)
    dfilm = torch.zeros_like(film)
    dbase = torch.zeros_like(base)
    dlogw = torch.zeros_like(logw)
    dpoints = torch.zeros_like(p)
    desc_obj = nat.StackDesc.from_buffer_copy(stack.desc)
    desc_obj.exchange = None
    desc_obj.nonfinite = None
    desc = ctypes.byref(desc_obj)
    P = nat.ptr
    n_total = float(B * N)
    keep = bool(stack.keep_activations) and int(lib.gwtf_keep_floats(desc, B, N)) > 0
    ybuf = torch.empty(int(lib.gwtf_keep_floats(desc, B, N)), device=dev) if keep else None

    def fwd():
        nat.check(lib.gwtf_fwd_all(desc, 1, P(params), P(bnbuf), P(film), P(p), P(base), P(logw), P(ubuf), P(ld), None,
                                   P(ybuf), P(mom), P(sum1), P(bstat), B, N, P(nll), None, _stream_ptr()), 'gwtf_fwd_all')

    def bwd():
        bsum.zero_()
        nat.check(lib.gwtf_bwd_all(desc, 1, P(params), P(bnbuf), P(film), P(p), P(base), P(logw), P(ubuf), P(ybuf), P(ld),
                                   P(mom), P(sum1), P(nll), P(dnll), P(bsum), P(gbuf), P(gs), P(dobuf), P(dparams),
                                   P(dfilm), P(dbase), P(dlogw), P(dpoints), B, N, _stream_ptr()), 'gwtf_bwd_all')

    def phase_loop(kind, phase):
        st = _stream_ptr()
        for l in (range(L - 1, -1, -1) if kind == 'fwd' else range(L)):
            if kind == 'fwd':
                nat.check(lib.gwtf_fwd_layer(desc, l, phase, 1, P(params), P(bnbuf), P(film), P(p), P(ubuf), P(ld),
                                             None, P(ybuf), P(mom), P(sum1), B, N, n_total, st), 'gwtf_fwd_layer')
            else:
                nat.check(lib.gwtf_bwd_layer(desc, l, phase, 1, P(params), P(bnbuf), P(film), P(p), P(ubuf), P(ybuf), P(mom),
                                             P(sum1), P(bsum), P(gbuf), P(gs), P(dobuf), P(dparams), P(dfilm), B, N,
                                             n_total, st), 'gwtf_bwd_layer')

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        torch.cuda.synchronize()
        best = []
        for _ in range(reps):
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best.append(e0.elapsed_time(e1))
        return sum(best) / len(best)

    out = {'fwd_ms': timed(fwd), 'bwd_ms': timed(bwd), 'kept': bool(keep), 'kept_bytes': 4 * ybuf.numel() if keep else 0}
    # one launch class at a time, ordinary launches: chaining a kernel behind ITSELF with programmatic dependent
    # launch (which the real sequence never does) makes the early CTAs of launch n+1 compete with launch n
    desc_obj.flags |= nat.FLAG_NO_PDL
    for name, kind, phase in (('fwd_stats', 'fwd', 0), ('fwd_apply', 'fwd', 1), ('bwd_d', 'bwd', 0), ('bwd_e', 'bwd', 1)):
        out[name + '_ms_per_launch'] = timed(lambda: phase_loop(kind, phase)) / L
    del ybuf
    return out


def rate(points_per_s, flops_per_pt, peaks, fma_peak, passes):
    """Roofline block of a whole pass: algorithmic TFLOP/s against the measured tensor peak for the precision tier
    (`passes` TF32 MMAs per product) and against the FP32 FMA pipe."""
    ach = points_per_s * flops_per_pt * 1e-12
    tensor_peak = peaks['bf16_tflops'] / 2.0 / passes
    return {'bound': 'tensor', 'achieved': ach, 'peak': tensor_peak, 'unit': 'TFLOP/s', 'frac': ach / tensor_peak,
            'peak_source': 'MEASURED_PEAKS bf16_tflops %.1f / 2 (TF32) / %d (TF32 MMAs per product), of %s' %
                           (peaks['bf16_tflops'], passes, peaks['source']),
            'fma_peak': fma_peak, 'frac_of_fma_peak': ach / fma_peak, 'flops_per_point': flops_per_pt, 'traffic': None}


def note(env, what):
    """Progress line on stderr: a stalled multi-rank run shows where it stalled."""
    if env.rank == 0:
        sys.stderr.write('[bench] %s\n' % what)
        sys.stderr.flush()


def side_workloads(env, args, peaks, fma_peak):
    """The other BASELINE configs, on every rank, aggregated over ranks."""
    from go_with_the_flows_b200 import _native as nat
    from go_with_the_flows_b200 import flowstack
    from go_with_the_flows_b200.flowstack import sample_mixture
    world, rank, dev = env.world, env.rank, env.dev
    N = args.points
    out = {}
    quick = args.quick

    note(env, 'side workload: eval_nll')
    # ---- eval-mode NLL, airplane model, 64 clouds per GPU: fp32-grade and single-pass TF32
    cfg, model = build_model('generative', dev)
    model.eval()
    model.mode = 'training'
    stack = model.flow_stack()
    K, L, Fd = stack.K, stack.L, stack.F
    fl_fwd, _ = flops_per_point(Fd, K, L)
    p, g = synthetic(64, N, cfg['g_latent_space_size'], seed_shift=7 + rank)
    p, g = p.to(dev), g.to(dev)
    ev = {}
    with torch.no_grad():
        for tier, prec, passes in (('fp32_grade_3xtf32', nat.PRECISION_3XTF32, 3), ('tf32', nat.PRECISION_TF32, 1)):
            stack.desc.eval_precision = prec
            ms = env.timed(lambda: model.decode(p, g, N), reps=3)
            pps = world * 64 * N / (ms * 1e-3)
            ev[tier] = {'points_per_s': pps, 'ms': ms, 'clouds_per_gpu': 64, 'points': N,
                        'roofline': rate(pps / world, fl_fwd, peaks, fma_peak, passes)}
        stack.desc.eval_precision = flowstack._DEFAULTS['eval_precision']
    out['eval_nll'] = ev

    note(env, 'side workload: sampling_c5')
    # ---- sampling sweep C5: airplane decoder, 256 latents per GPU
    fl_samp = L * 4 * (Fd * Fd + 3 * Fd)
    _, g2 = synthetic(256, 8, cfg['g_latent_space_size'], seed_shift=9 + rank)
    g2 = g2.to(dev)
    sweep = {}
    with torch.no_grad():
        logits = model.get_weights(g2)
        mu_b, lv_b = model.base_gaussian(g2)
        for tier, prec, passes in (('fp32_grade_3xtf32', nat.PRECISION_3XTF32, 3), ('tf32', nat.PRECISION_TF32, 1)):
            stack.desc.eval_precision = prec
            rows = {}
            for n_pts in ((2048, 16384) if quick else (2048, 16384, 131072, 1048576)):
                if tier == 'fp32_grade_3xtf32' and n_pts > 131072:
                    continue
                reps = 3 if n_pts <= 16384 else 1
                ms = env.timed(lambda: sample_mixture(stack, g2, mu_b, lv_b, logits, n_pts, 2026, rank), reps=reps, flush=False)
                pps = world * 256 * n_pts / (ms * 1e-3)
                rows[str(n_pts)] = {'points_per_s': pps, 'ms': ms, 'roofline': rate(pps / world, fl_samp, peaks, fma_peak, passes)}
                stack._keep_pool.clear()
                torch.cuda.empty_cache()
            sweep[tier] = rows
        stack.desc.eval_precision = flowstack._DEFAULTS['eval_precision']
        # B = 1 latency (the reference's evaluation loop samples shape by shape)
        ms = env.timed(lambda: sample_mixture(stack, g2[:1], mu_b[:1], lv_b[:1], logits[:1], N, 2026, rank), reps=5, flush=False)
        sweep['latency_1x%d_ms' % N] = ms
    out['sampling_c5'] = {'latents_per_gpu': 256, 'decoder': 'airplane (K=4,L=33,F=37)', 'tiers': sweep}
    del model, stack
    torch.cuda.empty_cache()

    note(env, 'side workload: sampling_c4')
    # ---- C4: config_SVR decoder (F=33, G=512, freevar), 256 x 2048 and 64 x 2500 (the config's cloud_size)
    cfg4, model4 = build_model('svr', dev)
    model4.eval()
    model4.mode = 'reconstruction'
    st4 = model4.flow_stack()
    fl4 = st4.L * 4 * (st4.F * st4.F + 3 * st4.F)
    c4 = {}
    with torch.no_grad():
        for tier, prec, passes in (('fp32_grade_3xtf32', nat.PRECISION_3XTF32, 3), ('tf32', nat.PRECISION_TF32, 1)):
            st4.desc.eval_precision = prec
            rows = {}
            for Bs, n_pts in ((256, 2048), (64, 2500)):
                _, g4 = synthetic(Bs, 8, cfg4['g_latent_space_size'], seed_shift=11 + rank)
                g4 = g4.to(dev)
                lg = model4.get_weights(g4)
                mb, lb = model4.base_gaussian(g4)
                ms = env.timed(lambda: sample_mixture(st4, g4, mb, lb, lg, n_pts, 2026, rank), reps=3, flush=False)
                pps = world * Bs * n_pts / (ms * 1e-3)
                rows['%dx%d' % (Bs, n_pts)] = {'points_per_s': pps, 'ms': ms,
                                               'roofline': rate(pps / world, fl4, peaks, fma_peak, passes)}
            c4[tier] = rows
    out['sampling_c4'] = {'decoder': 'config_SVR (K=4,L=33,F=33,G=512)', 'tiers': c4}
    del model4, st4
    torch.cuda.empty_cache()

    note(env, 'side workload: c3_step')
    # ---- C3: config_autoencoding train step, 128 clouds per GPU
    if not quick:
        cfg3, model3 = build_model('autoencoding', dev)
        model3.train()
        model3.mode = 'training'
        st3 = model3.flow_stack()
        B3 = 128
        p3, g3 = synthetic(B3, N, cfg3['g_latent_space_size'], seed_shift=13 + rank)
        p3, g3 = p3.to(dev), g3.to(dev)
        step3 = make_step(model3, world, N)
        ms = env.timed(lambda: step3(p3, g3), reps=3, warmup=2)
        _, fl3 = flops_per_point(st3.F, st3.K, st3.L)
        pps = world * B3 * N / (ms * 1e-3)
        out['c3_step'] = {'points_per_s': pps, 'ms_per_step': ms, 'clouds_per_gpu': B3,
                          'workload': 'config_autoencoding (K=4,L=33,F=33,G=512,freevar), train-mode fwd+bwd',
                          'roofline': rate(pps / world, fl3, peaks, fma_peak, 3)}
        del model3, st3, step3
        torch.cuda.empty_cache()

    note(env, 'side workload: small_step')
    # ---- small-batch train step (4 clouds x 2048 per GPU): launch-bound regime, 132 layer launches of ~10 us each
    if not quick:
        cfg4, model4 = build_model('generative', dev)
        model4.train()
        model4.mode = 'training'
        B4 = 4
        p4, g4 = synthetic(B4, N, cfg4['g_latent_space_size'], seed_shift=29 + rank)
        p4, g4 = p4.to(dev), g4.to(dev)
        step4 = make_step(model4, world, N)
        ms = env.timed(lambda: step4(p4, g4), reps=10, warmup=3, flush=False)
        out['small_step'] = {'ms_per_step': ms, 'points_per_s': world * B4 * N / (ms * 1e-3), 'clouds_per_gpu': B4,
                             'workload': 'airplane model, train-mode fwd+bwd, 4 clouds x 2048 per GPU (CUDA events '
                                         'around the whole step incl. the PyTorch launches of the FiLM nets)'}
        del model4, step4
        torch.cuda.empty_cache()

    note(env, 'side workload: strong')
    # ---- strong scaling at the reference's semantics: 64 clouds in total (train_ae.py:77-78)
    if world > 1 and not quick:
        cfgs, models = build_model('generative', dev)
        models.train()
        models.mode = 'training'
        Bs = max(1, 64 // world)
        ps, gs = synthetic(64, N, cfgs['g_latent_space_size'])
        ps, gs = ps[rank * Bs:(rank + 1) * Bs].contiguous().to(dev), gs[rank * Bs:(rank + 1) * Bs].contiguous().to(dev)
        steps_ = make_step(models, world, N)
        ms = env.timed(lambda: steps_(ps, gs), reps=5, warmup=3)
        _, fls = flops_per_point(models.flow_stack().F, models.flow_stack().K, models.flow_stack().L)
        pps = world * Bs * N / (ms * 1e-3)
        out['strong'] = {'points_per_s': pps, 'ms_per_step': ms, 'clouds_total': Bs * world, 'clouds_per_gpu': Bs,
                         'scaling': 'strong', 'roofline': rate(pps / world, fls, peaks, fma_peak, 3)}
        del models, steps_
        torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='generative', choices=['generative', 'autoencoding', 'svr'])
    ap.add_argument('--batch', type=int, default=64, help='clouds per GPU')
    ap.add_argument('--points', type=int, default=2048)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='headline only (skip the side workloads)')
    ap.add_argument('--quick', action='store_true', help='shorter side workloads')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    # stdout carries exactly ONE line, the JSON: anything a library prints there (NCCL's version banner ...) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback for the flow kernels)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    args.warmup = max(args.warmup, 3)
    env = Env(world, rank, dev)

    from go_with_the_flows_b200 import _native as nat
    cfg, model = build_model(args.config, dev)
    model.train()
    model.mode = 'training'
    stack = model.flow_stack()
    B, N, G = args.batch, args.points, cfg['g_latent_space_size']
    p_host, g_host = synthetic(B, N, G, seed_shift=rank)
    p_host, g_host = p_host.pin_memory(), g_host.pin_memory()
    p_dev, g_dev = p_host.to(dev), g_host.to(dev)
    loss_host = torch.zeros((), pin_memory=True)
    step = make_step(model, world, N)

    for _ in range(args.warmup):
        step(p_dev, g_dev)
    env.barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0 and os.environ.get('GWTF_BENCH_NO_SAMPLER') != '1':
        sampler.start()
    note(env, 'timed steps')
    # ---- device-resident timing: K steps, each bracketed by CUDA events, L2 flushed in between
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    env.barrier()
    torch.cuda.profiler.start()          # (`ncu --profile-from-start off` then lists exactly the timed steps; a no-op otherwise)
    for e0, e1 in evs:
        env.flush_buf.fill_(0.0)
        e0.record()
        step(p_dev, g_dev)
        e1.record()
    env.barrier()
    torch.cuda.profiler.stop()
    dev_ms = env.max_over_ranks(sum(e0.elapsed_time(e1) for e0, e1 in evs))
    note(env, 'end-to-end steps')
    # ---- end to end: pinned host inputs -> public API -> loss on the host, wall clock
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        p = p_host.to(dev, non_blocking=True)
        g = g_host.to(dev, non_blocking=True)
        pnll = step(p, g)
        loss_host.copy_(pnll.detach(), non_blocking=True)
        torch.cuda.synchronize()
    env.barrier()
    e2e_ms = env.max_over_ranks((time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop() if rank == 0 else None
    loss_value = float(loss_host)
    nonfinite = model.nonfinite_points()

    peaks = measured_peaks()
    fma = ctypes.c_double(0.0)
    nat.check(nat.lib().gwtf_fma_peak_tflops(20000, ctypes.byref(fma), None), 'gwtf_fma_peak_tflops')
    fma_peak = fma.value
    extra = None if args.no_extra else side_workloads(env, args, peaks, fma_peak)

    if rank == 0:
        K, L, Fd = stack.K, stack.L, stack.F
        fl_fwd, fl_step = flops_per_point(Fd, K, L)
        pts_per_step = world * B * N
        ms_per_step = dev_ms / args.steps
        value = pts_per_step / (ms_per_step * 1e-3)
        e2e_value = pts_per_step / (e2e_ms / args.steps * 1e-3)
        kt = kernel_only_times(model, p_dev, g_dev, reps=3)
        mma_peak = ctypes.c_double(0.0)
        nat.check(nat.lib().gwtf_mma_peak_tflops(4000, ctypes.byref(mma_peak), None), 'gwtf_mma_peak_tflops')
        fwd_eng, bwd_eng = stack.fwd_engine(), stack.bwd_engine()
        kept = bool(kt['kept'])
        kernel_ms = kt['fwd_ms'] + kt['bwd_ms']
        achieved = B * N * fl_step / (kernel_ms * 1e-3) * 1e-12
        per_pc_layer = 4 * (Fd * Fd + 3 * Fd)      # fwd FLOPs per point, component, layer
        launch_units = K * B * N
        tensor_peak = peaks['bf16_tflops'] / 2.0 / 3.0          # dense TF32 = bf16 / 2; fp32-grade 3xTF32 = / 3
        # algorithmic share of each launch class (recomputation earns nothing)
        alg = {'fwd_stats': 0.0, 'fwd_apply': per_pc_layer * launch_units, 'bwd_d': 4 * Fd * 3 * launch_units,
               'bwd_e': (2 * per_pc_layer - 4 * Fd * 3) * launch_units}
        names = {'fwd_stats': 'k_fwd_layer_tcp<..,0>', 'fwd_apply': 'k_fwd_layer_tcp<..,1>',
                 'bwd_d': 'k_bwd_layer_tc<..,0>' if bwd_eng == nat.ENGINE_TC else 'k_bwd_layer_d_mma',
                 'bwd_e': 'k_bwd_layer_tc<..,1>' if bwd_eng == nat.ENGINE_TC else 'k_bwd_layer_e_mma'}
        kernels = {}
        for name in ('fwd_stats', 'fwd_apply', 'bwd_d', 'bwd_e'):
            ms = kt[name + '_ms_per_launch']
            tf = alg[name] / (ms * 1e-3) * 1e-12
            kernels[name] = {'kernel': names[name], 'ms_per_launch': ms, 'share_of_step_kernels': ms * L / kernel_ms,
                             'algorithmic_tflops': tf, 'frac': tf / tensor_peak, 'frac_of_fma_peak': tf / fma_peak}
        dom_name = max(('fwd_stats', 'fwd_apply', 'bwd_d', 'bwd_e'), key=lambda n: kt[n + '_ms_per_launch'])
        dom = kernels[dom_name]
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'r02_ncu_traffic.json')
        if os.path.exists(tpath) and args.config == 'generative' and B == 64 and N == 2048:
            with open(tpath) as fh:       # dram bytes per launch of the dominant kernel from the committed ncu capture
                traffic = json.load(fh).get(dom['kernel'])
        line = {
            'metric': 'points/sec fwd+bwd mixture-flow NLL', 'value': value, 'unit': 'points/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(args), 'clouds_per_gpu': B, 'points_per_cloud': N,
                       'parallelism': 'dp%d (shapes sharded, SyncBN statistics + gradient all-reduce)' % world,
                       'l2': 'flushed between timed steps (256 MiB write outside the event bracket)',
                       'bn': 'train mode (batch statistics)', 'loss': loss_value, 'nonfinite_points': nonfinite},
            'e2e': {'value': e2e_value, 'unit': 'points/s', 'h2d_bytes_per_step': world * (p_host.numel() + g_host.numel()) * 4,
                    'd2h_bytes_per_step': world * 4, 'ms_per_step': e2e_ms / args.steps},
            # our kernels per step: 4 layer phases x L, moments/bstat/nll/seed/2 finish, + one exchange kernel per phase at N > 1
            'gpu_launches': args.steps * (4 * L + 6 + (4 * L if world > 1 else 0)),
            'roofline': {'bound': 'tensor',
                         'kernel': '%s (dominant: %.0f %% of the step kernels; per launch = one coupling layer, all K '
                                   'components)' % (dom['kernel'], 100 * dom['share_of_step_kernels']),
                         'achieved': dom['algorithmic_tflops'], 'peak': tensor_peak, 'unit': 'TFLOP/s',
                         'frac': dom['algorithmic_tflops'] / tensor_peak, 'traffic': traffic,
                         'peak_source': 'MEASURED_PEAKS.json bf16_tflops %.1f (burst, of %s) / 2 (dense TF32 = half the bf16 '
                                        'rate) / 3 (fp32-grade 3xTF32: three MMAs per product)' %
                                        (peaks['bf16_tflops'], peaks['source']),
                         'fma_peak': fma_peak, 'frac_of_fma_peak': dom['algorithmic_tflops'] / fma_peak,
                         'diagnostic_mma_sync_tf32_peak': mma_peak.value,
                         'step': {'kernel_ms': kernel_ms, 'fwd_ms': kt['fwd_ms'], 'bwd_ms': kt['bwd_ms'],
                                  'achieved': achieved, 'frac': achieved / tensor_peak,
                                  'frac_of_fma_peak': achieved / fma_peak, 'flops_per_point': fl_step,
                                  'target': '>= 0.50 of the FP32-FMA roofline (north star)'},
                         'kernels': kernels},
            'clocks': clocks,
            'engine': 'forward %s, backward %s%s' % (nat.ENGINE_NAMES[fwd_eng], nat.ENGINE_NAMES[bwd_eng],
                                                     '; sd1 outputs kept for backward (%.1f GB)' % (kt['kept_bytes'] / 1e9)
                                                     if kept else '; backward recomputes from the 12-byte layer inputs'),
            'extra': extra,
        }
        if not args.no_cpu_baseline and world == 1:
            cb, cn = 4, N
            sec, cores, kind = cpu_step_time(args.config, cb, cn, 2, 1)
            line['cpu_baseline'] = {'value': cb * cn / sec, 'unit': 'points/s', 'cores': cores, 'kind': kind,
                                    'sample': 'same model, train-mode fwd+bwd, %d clouds x %d points, 2 timed steps '
                                              'after 1 warm-up (%.2f s/step)' % (cb, cn, sec)}
        os.write(json_fd, (json.dumps(line) + '\n').encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
