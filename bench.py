#!/usr/bin/env python
"""Benchmark of the mixture-of-flows hot path (BASELINE.json metric: points/sec fwd+bwd mixture NLL).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W     # the reference algorithm on host cores

A step = one train-mode (batch-statistics BatchNorm) forward + backward of the K-component
mixture-of-flows NLL over one synthetic batch: `Flow_Mixture_Model.decode` + `FlowMixtureNLL` +
`.backward()`, i.e. everything between the shape latent / point cloud and the gradients of every
decoder parameter, the latent, the base Gaussian and the mixture weights.  Workload = BASELINE.json
configs[1]: config_generative_modeling_airplane.yaml model (K=4, 33 coupling layers, F=37, G=128),
64 clouds x 2048 points per GPU, random init (seed 0), synthetic clouds N(0, 0.2^2), latents N(0, 0.5^2).

Prints ONE JSON line (rank 0).  `value` times the step with inputs resident in HBM; `e2e` times the
same step through the public module API from pinned host buffers (H2D of points + latents, D2H of
the loss inside the timed region).  `roofline` compares the step's own kernels (CUDA-event timed, no
Python in between) with the FP32 FMA pipe, which is what binds this path (~49 kFLOP
per byte); the FMA peak is measured live by an FFMA probe kernel on the same GPU.
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


def oracle_step_time(cfg_name, B, N, steps, warmup):
    """The reference algorithm (CPU restatement in oracle/, same ATen ops as lib/networks) timed on
    the host cores: train-mode forward + backward of the mixture NLL on a bounded sample."""
    from oracle import flow_oracle as fo
    torch.set_num_threads(os.cpu_count() or 1)
    cfg, model = build_model(cfg_name, 'cpu')
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for k, v in sd.items():
        if v.is_floating_point() and k.startswith(('pc_decoder', 'p_prior', 'mixture_weights')) \
                and 'running' not in k and not k.endswith('eps'):
            v.requires_grad_(True)
    p, g = synthetic(B, N, cfg['g_latent_space_size'])
    g.requires_grad_(True)

    def step():
        for v in sd.values():
            v.grad = None
        out = fo.mixture_nll(p, g, sd, base_type=cfg['p_decoder_base_type'], weights_type=cfg['weights_type'],
                             training=True, base_var=cfg['p_decoder_base_var'])
        out['pnll'].backward()
        return float(out['pnll'])

    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return sum(times) / len(times), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    B, N = 4, args.points
    steps = max(1, min(args.steps, 3))
    warmup = max(1, min(args.warmup, 1))
    sec, cores = oracle_step_time(args.config, B, N, steps, warmup)
    value = B * N / sec
    sample = 'train-mode fwd+bwd mixture NLL, %d clouds x %d points per step (autograd holds ~1 GB/shape), ' \
             '%d timed steps after %d warm-up' % (B, N, steps, warmup)
    line = {
        'impl': 'reference', 'metric': 'points/sec fwd+bwd mixture-flow NLL', 'value': value, 'unit': 'points/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args), 'cpu_sample_clouds': B},
        'cpu_baseline': {'value': value, 'unit': 'points/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'points/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def workload_name(args):
    return 'C2: config_generative_modeling_airplane model (K=4,L=33,F=37,G=128), train-mode fwd+bwd mixture NLL, ' \
           '%d clouds x %d points per GPU' % (args.batch, args.points) if args.config == 'generative' else \
           '%s config, train-mode fwd+bwd mixture NLL, %d clouds x %d points per GPU' % (args.config, args.batch,
                                                                                       args.points)


def kernel_only_times(model, p, g, reps, keep=None):
    """CUDA-event time of the step's own kernels with nothing in between: forward driver
    (moments + 2 phases x L layers + bstat + nll) and backward driver (seed + 2 phases x L + finish),
    plus each phase class timed over the L layers."""
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
    desc = ctypes.byref(stack.desc)
    P = nat.ptr
    n_total = float(B * N)
    # kept activations follow the product default (FlowStack.keep_activations: on for the mma engine)
    if keep is None:
        keep = stack.keep_activations if stack.keep_activations is not None else lib.gwtf_engine() != 0
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
    prev_pdl = lib.gwtf_set_pdl(0)
    for name, kind, phase in (('fwd_stats', 'fwd', 0), ('fwd_apply', 'fwd', 1), ('bwd_d', 'bwd', 0), ('bwd_e', 'bwd', 1)):
        out[name + '_ms_per_launch'] = timed(lambda: phase_loop(kind, phase)) / L
    lib.gwtf_set_pdl(prev_pdl)
    del ybuf
    return out


def nat_engine():
    from go_with_the_flows_b200 import _native as nat
    return int(nat.lib().gwtf_engine())


def side_workloads(model, cfg, dev, N):
    """Other hot-path entry points (not the headline): fused eval-mode NLL (BASELINE configs[0] shape
    at 64 clouds) and sampling (configs[3]/[4] shape: 256 latents x N points).  CUDA-event timed."""
    from go_with_the_flows_b200.flowstack import sample_mixture
    out = {}
    G = cfg['g_latent_space_size']
    was_training = model.training
    model.eval()
    try:
        with torch.no_grad():
            p, g = synthetic(64, N, G, seed_shift=7)
            p, g = p.to(dev), g.to(dev)

            def timed(fn, reps=3):
                fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / reps

            ms = timed(lambda: model.decode(p, g, N))
            out['eval_nll_fused'] = {'points_per_s': 64 * N / (ms * 1e-3), 'ms': ms, 'clouds': 64, 'points': N,
                                     'kernel': 'per-layer tensor-core kernels, eval-mode BN (gwtf_nll_fwd_eval_layers)' if nat_engine() != 0 else 'k_nll_eval (one launch, fp32 FMA)'}
            _, g2 = synthetic(256, N, G, seed_shift=9)
            g2 = g2.to(dev)
            stack = model.flow_stack()
            logits = model.get_weights(g2)
            mu_b, lv_b = model.base_gaussian(g2)
            ms = timed(lambda: sample_mixture(stack, g2, mu_b, lv_b, logits, N, 2026, 0))
            out['sampling'] = {'points_per_s': 256 * N / (ms * 1e-3), 'ms': ms, 'latents': 256, 'points': N,
                               'kernel': 'Philox draws, points regrouped by component, per-layer tensor-core kernels in direct mode (gwtf_sample_layers)' if nat_engine() != 0 else 'k_sample (Philox draws + direct stacks, fp32 FMA)'}
    finally:
        model.train(was_training)
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
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback for the flow kernels)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    args.warmup = max(args.warmup, 3)

    from go_with_the_flows_b200 import _native as nat
    from go_with_the_flows_b200.networks.losses import FlowMixtureNLL
    cfg, model = build_model(args.config, dev)
    model.train()
    model.mode = 'training'
    loss_fn = FlowMixtureNLL()
    stack = model.flow_stack()
    B, N, G = args.batch, args.points, cfg['g_latent_space_size']
    p_host, g_host = synthetic(B, N, G, seed_shift=rank)
    p_host, g_host = p_host.pin_memory(), g_host.pin_memory()
    p_dev, g_dev = p_host.to(dev), g_host.to(dev)
    loss_host = torch.zeros((), pin_memory=True)
    params_all = [q for q in model.parameters()]
    stack.prepare()
    owned = set()
    for name in stack.grad_masters:          # decoder tensors: averaged across ranks inside backward
        owned.update(id(q) for q in stack.masters[name].params)
    params_other = [q for q in params_all if id(q) not in owned]
    flush_buf = torch.empty(256 * 1024 * 1024 // 4, device=dev)

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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(p_dev, g_dev)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0 and os.environ.get('GWTF_BENCH_NO_SAMPLER') != '1':
        sampler.start()
    # ---- device-resident timing: K steps, each bracketed by CUDA events, L2 flushed in between
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for e0, e1 in evs:
        flush_buf.fill_(0.0)
        e0.record()
        step(p_dev, g_dev)
        e1.record()
    barrier()
    dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    # ---- end to end: pinned host inputs -> public API -> loss on the host, wall clock
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        p = p_host.to(dev, non_blocking=True)
        g = g_host.to(dev, non_blocking=True)
        pnll = step(p, g)
        loss_host.copy_(pnll.detach(), non_blocking=True)
        torch.cuda.synchronize()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    loss_value = float(loss_host)

    t = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        K, L, Fd = stack.K, stack.L, stack.F
        fl_fwd, fl_step = flops_per_point(Fd, K, L)
        pts_per_step = world * B * N
        ms_per_step = dev_ms / args.steps
        value = pts_per_step / (ms_per_step * 1e-3)
        e2e_value = pts_per_step / (e2e_ms / args.steps * 1e-3)
        kt = kernel_only_times(model, p_dev, g_dev, reps=3)
        peak = ctypes.c_double(0.0)
        nat.check(nat.lib().gwtf_fma_peak_tflops(20000, ctypes.byref(peak), None), 'gwtf_fma_peak_tflops')
        mma_peak = ctypes.c_double(0.0)
        nat.check(nat.lib().gwtf_mma_peak_tflops(4000, ctypes.byref(mma_peak), None), 'gwtf_mma_peak_tflops')
        engine = int(nat.lib().gwtf_engine())
        kept = bool(kt['kept'])
        kernel_ms = kt['fwd_ms'] + kt['bwd_ms']
        achieved = B * N * fl_step / (kernel_ms * 1e-3) * 1e-12
        per_pc_layer = 4 * (Fd * Fd + 3 * Fd)      # fwd FLOPs per point, component, layer
        launch_units = K * B * N
        # algorithmic share of each launch class (recomputation earns nothing)
        alg = {'fwd_stats': 0.0, 'fwd_apply': per_pc_layer * launch_units, 'bwd_d': 4 * Fd * 3 * launch_units,
               'bwd_e': (2 * per_pc_layer - 4 * Fd * 3) * launch_units}
        # F x F contractions each launch class executes per point, component and net
        executed = {'fwd_stats': 1, 'fwd_apply': 1, 'bwd_d': 0 if kept else 1, 'bwd_e': 2 if kept else 3}
        kernels = {}
        for name in ('fwd_stats', 'fwd_apply', 'bwd_d', 'bwd_e'):
            ms = kt[name + '_ms_per_launch']
            kernels[name] = {'ms_per_launch': ms, 'algorithmic_tflops': alg[name] / (ms * 1e-3) * 1e-12,
                             'executed_contractions': executed[name]}
        # tensor-pipe work of the dominant kernel: m16n8k8 tf32 MMAs per 16-point tile and net (F padded to 8,
        # the dW1 output to 16 rows), three per product (3xTF32)
        f8, f16 = (Fd + 7) // 8, (Fd + 15) // 16
        mma_per_tile = 3 * ((executed['bwd_e'] - 1) * f8 * f8 + 2 * f16 * f8)
        e_tensor_tflops = mma_per_tile * 2048.0 * (launch_units / 16.0) * 2 / (kernels['bwd_e']['ms_per_launch'] * 1e-3) * 1e-12
        traffic = None
        tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'r01_ncu_traffic.json')
        if os.path.exists(tpath) and args.config == 'generative' and B == 64 and N == 2048:
            with open(tpath) as fh:       # dram bytes per launch of the dominant kernel from the committed ncu capture
                traffic = json.load(fh).get('k_bwd_layer_e_mma', {}).get('kept' if kept else 'recompute')
        dom = kernels['bwd_e']
        line = {
            'metric': 'points/sec fwd+bwd mixture-flow NLL', 'value': value, 'unit': 'points/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(args), 'clouds_per_gpu': B, 'points_per_cloud': N,
                       'parallelism': 'dp%d (shapes sharded, SyncBN statistics + gradient all-reduce)' % world,
                       'l2': 'flushed between timed steps (256 MiB write outside the event bracket)',
                       'bn': 'train mode (batch statistics)', 'loss': loss_value},
            'e2e': {'value': e2e_value, 'unit': 'points/s', 'h2d_bytes_per_step': world * (p_host.numel() + g_host.numel()) * 4,
                    'd2h_bytes_per_step': world * 4, 'ms_per_step': e2e_ms / args.steps},
            # our kernels per step: 4 layer phases x L, moments/bstat/nll/seed/finish, + one exchange kernel per phase at N > 1
            'gpu_launches': args.steps * (4 * L + 6 + (4 * L if world > 1 else 0)),
            'roofline': {'bound': 'tensor',
                         'kernel': 'k_bwd_layer_e_mma (dominant; per launch = one coupling layer, all K components)',
                         'achieved': dom['algorithmic_tflops'], 'peak': mma_peak.value / 3.0, 'unit': 'TFLOP/s',
                         'frac': dom['algorithmic_tflops'] / (mma_peak.value / 3.0), 'traffic': traffic,
                         'peak_source': 'fp32-grade contractions run as 3xTF32 on mma.sync: peak = dense TF32 mma.sync '
                                        'throughput measured live by gwtf_mma_peak_tflops (%.1f TFLOP/s) / 3; '
                                        'MEASURED_PEAKS.json only holds the bf16 cuBLAS figure, which no fp32-parity '
                                        'path can use' % mma_peak.value,
                         'tensor_pipe': {'executed_tf32_tflops': e_tensor_tflops, 'mma_sync_tf32_peak': mma_peak.value,
                                         'frac': e_tensor_tflops / mma_peak.value},
                         'fma_peak': peak.value, 'frac_of_fma_peak': dom['algorithmic_tflops'] / peak.value,
                         'step': {'kernel_ms': kernel_ms, 'fwd_ms': kt['fwd_ms'], 'bwd_ms': kt['bwd_ms'],
                                  'achieved': achieved, 'frac': achieved / (mma_peak.value / 3.0),
                                  'frac_of_fma_peak': achieved / peak.value, 'flops_per_point': fl_step},
                         'kernels': kernels},
            'clocks': clocks,
            'engine': {0: 'fp32 FMA kernels',
                       1: 'tcgen05 3xTF32 forward (one tile per CTA) + mma.sync 3xTF32 backward',
                       2: 'tcgen05 3xTF32 persistent warp-specialised forward + mma.sync 3xTF32 backward',
                       3: 'mma.sync 3xTF32 forward and backward'}[engine] +
                      ('; sd1 outputs kept for backward (%.1f GB)' % (kt['kept_bytes'] / 1e9) if kept else '; backward recomputes h1') +
                      '; fp32 FMA fused-eval / sampling kernels',
            'extra': side_workloads(model, cfg, dev, N),
        }
        if not args.no_cpu_baseline and world == 1:
            cb, cn = 4, N
            sec, cores = oracle_step_time(args.config, cb, cn, 2, 1)
            line['cpu_baseline'] = {'value': cb * cn / sec, 'unit': 'points/s', 'cores': cores, 'kind': 'port',
                                    'sample': 'same model, train-mode fwd+bwd, %d clouds x %d points, 2 timed steps '
                                              'after 1 warm-up (%.2f s/step)' % (cb, cn, sec)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
