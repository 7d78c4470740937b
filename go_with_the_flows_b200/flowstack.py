"""Host side of the fused flow stack: parameter packing, batched FiLM conditioning nets and the
autograd Functions that call the C ABI (libgwtf.so).

A `FlowStack` views K lists of `CondRealNVPFlow3D` modules (the coupling layers of K mixture
components, in DIRECT order) as one (K, L) grid:

  * `pack_params()`  -> (K, L, rec_stride) fp32, autograd-connected to the module parameters
                        (layout documented in include/gwtf.h);
  * `pack_bn()`      -> (K, L, 8F) running statistics of the point-wise BatchNorms;
  * `film(g)`        -> (B, K, L, 2, 2, F): the 4 conditioning MLPs of every layer
                        (flows.py:33-45,68-80) evaluated for the whole grid as two batched GEMMs
                        in PyTorch -- latent-sized work that cuBLAS serves;
  * `nll_pass(p, g)` -> base-space samples z and per-dim log-det sums, differentiable, through
                        the phased (batch-statistics) or fused (running-statistics) CUDA kernels.

Nothing here computes the per-point flow in PyTorch: without libgwtf.so every entry raises.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import _native as nat

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def _stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return 1


def _env_int(name, default):
    import os
    v = os.environ.get(name)
    try:
        return int(v) if v not in (None, '') else default
    except ValueError:
        return default


# Execution options new FlowStacks start with (they live in each stack's descriptor, not in the library):
#   engine          GWTF_ENGINE = 4 tcgen05 forward + backward (default) | 2 tcgen05 forward + mma.sync backward |
#                   3 mma.sync | 0 FP32 FMA
#   pdl             GWTF_PDL = 1 programmatic dependent launch of the layer kernels (default) | 0
#   eval_precision  GWTF_EVAL_PRECISION = 0 fp32-grade 3xTF32 | 1 single-pass TF32 for the no-grad eval-mode NLL and
#                   the sampling pass
_DEFAULTS = {'engine': _env_int('GWTF_ENGINE', nat.ENGINE_TC), 'pdl': _env_int('GWTF_PDL', 1),
             'eval_precision': _env_int('GWTF_EVAL_PRECISION', nat.PRECISION_3XTF32)}


def set_default(name, value):
    """Set an execution option for FlowStacks created from now on; returns the previous value."""
    prev = _DEFAULTS[name]
    _DEFAULTS[name] = value
    return prev


_PEER = {'state': None, 'keep': None, 'handle': None}     # process-wide peer-memory statistic exchange (see gwtf.h)


def peer_exchange(slot_doubles, dev):
    """Create the NVLink peer-memory statistic exchange (csrc/gwtf_exchange.cuh) once per process group: a
    symmetric-memory buffer per rank holding the flag array and the double-buffered receive slots.
    Returns the exchange handle (int) when usable, None when the ranks should fall back to NCCL all-reduces
    of the statistic arrays (CPU/gloo groups, GWTF_PEER_EXCHANGE=0, no symmetric memory)."""
    import os
    st = _PEER['state']
    if _PEER.get('group') != (dist.get_world_size(), dist.get_rank()):
        if _PEER['handle']:
            nat.lib().gwtf_exchange_destroy(ctypes.c_void_p(_PEER['handle']))
        st = _PEER['state'] = None                  # a new process group: attach again
        _PEER['handle'] = None
        _PEER['group'] = (dist.get_world_size(), dist.get_rank())
    if st is not None and (st is False or st >= slot_doubles):
        return _PEER['handle'] if st else None
    if os.environ.get('GWTF_PEER_EXCHANGE', '1') == '0' or dist.get_backend() != 'nccl':
        _PEER['state'] = False
        return None
    world, rank = dist.get_world_size(), dist.get_rank()
    ok = torch.ones(1, device=dev)
    try:
        import torch.distributed._symmetric_memory as symm
        slot = max(int(slot_doubles), 2048)
        head = 32                                   # doubles reserved for the flag array (>= world uint64)
        buf = symm.empty(head + 4 * world * slot, dtype=torch.float64, device=dev)   # [2][world][slot] 16-byte cells
        buf.zero_()
        hdl = symm.rendezvous(buf, dist.group.WORLD)
        ptrs = [int(q) for q in hdl.buffer_ptrs]
        torch.cuda.synchronize(dev)
        hdl.barrier()
    except Exception:                               # any rank failing means every rank falls back
        ok.zero_()
        ptrs = None
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if float(ok.item()) < 1.0:
        _PEER['state'] = False
        return None
    arr = ctypes.c_void_p * world
    recv = arr(*[ctypes.c_void_p(q + head * 8) for q in ptrs])
    flags = arr(*[ctypes.c_void_p(q) for q in ptrs])
    out = ctypes.c_void_p()
    if _PEER['handle']:
        nat.lib().gwtf_exchange_destroy(ctypes.c_void_p(_PEER['handle']))
    timeout = float(os.environ.get('GWTF_EXCHANGE_TIMEOUT_S', '600'))
    nat.check(nat.lib().gwtf_exchange_create(rank, world, recv, flags, slot, timeout, ctypes.byref(out)),
              'gwtf_exchange_create')
    _PEER['state'] = slot
    _PEER['keep'] = (buf, hdl)
    _PEER['handle'] = out.value
    return out.value


class _LayerRef:
    """Direct references to the tensors of one coupling layer, in record order."""
    __slots__ = ('module', 'warp', 'point', 'cond', 'bn_point')

    def __init__(self, m):
        self.module = m
        self.warp = list(m.warp_inds)
        self.point = []      # parameters in record order
        self.bn_point = []   # (bn0, bn1) per net
        self.cond = []       # per net, per (w, b): (lin0, bn, lin1)
        for X in ('mu', 'logvar'):
            t0 = getattr(m, 'T_%s_0' % X)
            t1 = getattr(m, 'T_%s_1' % X)
            sd0, bn0, sd1, bn1 = t0[0], t0[1], t0[3], t0[4]
            sd2 = t1[1]
            self.point += [sd0.weight, bn0.weight, bn0.bias, sd1.weight, sd2.weight, sd2.bias]
            self.bn_point.append((bn0, bn1))
            for which in ('w', 'b'):
                seq = getattr(m, 'T_%s_0_cond_%s' % (X, which))
                self.cond.append((seq[0], seq[1], seq[3]))


class _Master:
    """One flat leaf tensor that backs many module parameters (or buffers) as views."""

    def __init__(self, name, shape, dtype=torch.float32):
        self.name = name
        self.shape = shape
        self.dtype = dtype
        self.tensor = None
        self.members = []      # (owner module, attribute name, is_param, offset, shape)
        self.grad_views = None
        self.grad_buf = None   # persistent gradient storage the per-parameter .grad views point into
        self.params = None     # Parameter objects of the members (grad masters only)

    def add(self, owner, attr, is_param, offset, shape):
        self.members.append((owner, attr, is_param, offset, tuple(shape)))

    def current(self, m):
        owner, attr, is_param, _, _ = m
        return owner._parameters[attr] if is_param else owner._buffers[attr]


class FlowStack:
    """K x L grid of coupling layers whose tensors live in a few flat 'master' tensors.

    The module Parameters / buffers keep their identity, names and shapes (state_dict, optimizers
    and checkpoints see nothing new) but their storage is a view into a master laid out the way
    the kernels read it, so packing is free and autograd runs on 6 leaves instead of ~4,800:

      point   (K, L, rec_stride)   sd0 / bn0 affine / sd1 / sd2 of both nets   (kernel `params`)
      bn      (K, L, 8F)           running stats of the point-wise BatchNorms   (kernel `bnbuf`)
      c_w0    (C, F, G)            first Linear of the C = K*L*4 FiLM nets
      c_bnw, c_bnb (C, F)          their BatchNorm affine;  c_rm, c_rv (C, F) running stats
      c_w1    (C, F, F), c_b1 (C, F)   second Linear
      nbt     (n,) int64           every num_batches_tracked counter

    Gradients: autograd accumulates into `master.grad`; after each backward the Parameters' `.grad`
    are (re)attached as views of it.  Because DistributedDataParallel never sees the masters, the
    decoder gradients are averaged across ranks here (one all-reduce per master, issued from the
    backward pass) whenever a process group is initialised -- the reference's bucketed DDP
    all-reduce (train_ae.py:153) for these tensors, as a single flat buffer.
    """

    def __init__(self, components, own_storage=True):
        """components: K lists of L CondRealNVPFlow3D modules (direct order).  own_storage=False
        (the per-module list API) leaves the module tensors where they are and gathers them per
        call instead -- only one stack may own the storage of a given set of modules."""
        self.flat = own_storage
        self._components = components
        self.K = len(components)
        self.L = len(components[0])
        assert all(len(c) == self.L for c in components)
        self.layers = [[_LayerRef(m) for m in comp] for comp in components]
        first = components[0][0]
        self.F = first.f_n_features
        self.G = first.g_n_features
        self.desc = nat.StackDesc()
        self.desc.n_components = self.K
        self.desc.n_layers = self.L
        self.desc.n_features = self.F
        self.rec_stride = 4 * ((2 * (self.F * self.F + 5 * self.F + 2) + 3) // 4)
        self.desc.rec_stride = self.rec_stride
        for l in range(self.L):
            warp = self.layers[0][l].warp
            for j in range(self.K):
                if self.layers[j][l].warp != warp:
                    raise ValueError('all components must share the warp pattern of each layer')
            mask = 0
            for d in warp:
                mask |= 1 << d
            self.desc.warp_mask[l] = mask
        self.desc.engine = _DEFAULTS['engine']
        self.desc.flags = 0 if _DEFAULTS['pdl'] else nat.FLAG_NO_PDL
        self.desc.eval_precision = _DEFAULTS['eval_precision']
        self.desc.exchange = None
        self.desc.nonfinite = None
        self._nonfinite = None       # device counter of non-finite per-point NLLs (training.py:43-46 stop condition)
        self.sync_gradients = True
        self._n_total_cache = {}
        # keep the sd1 output of every layer/net from the forward apply pass so that the mma.sync backward phases
        # skip one F x F contraction each (320 B per point/component/layer at F=37: 5.5 GB for 64 x 2048
        # points).  Off by default: the backward recomputes from the 12-byte layer inputs (the tcgen05 backward
        # always does); True trades 20 GB of HBM traffic per step for ~6 % on the mma.sync engines.
        self.keep_activations = False
        self._keep_pool = {}
        self.C = self.K * self.L * 4
        if self.flat:
            self._build_layout()

    # ------------------------------------------------------------------ flat storage
    def _build_layout(self):
        K, L, Fd, G = self.K, self.L, self.F, self.G
        C = K * L * 4
        self.C = C
        M = {
            'point': _Master('point', (K, L, self.rec_stride)),
            'bn': _Master('bn', (K, L, 8 * Fd)),
            'c_w0': _Master('c_w0', (C, Fd, G)),
            'c_bnw': _Master('c_bnw', (C, Fd)),
            'c_bnb': _Master('c_bnb', (C, Fd)),
            'c_rm': _Master('c_rm', (C, Fd)),
            'c_rv': _Master('c_rv', (C, Fd)),
            'c_w1': _Master('c_w1', (C, Fd, Fd)),
            'c_b1': _Master('c_b1', (C, Fd)),
            'nbt': _Master('nbt', (K * L * 8,), torch.int64),
        }
        self.masters = M
        self.grad_masters = ('point', 'c_w0', 'c_bnw', 'c_bnb', 'c_w1', 'c_b1')
        c = 0
        n = 0
        for j in range(K):
            for l in range(L):
                ref = self.layers[j][l]
                m = ref.module
                off = (j * L + l) * self.rec_stride
                for X in ('mu', 'logvar'):
                    t0, t1 = getattr(m, 'T_%s_0' % X), getattr(m, 'T_%s_1' % X)
                    sd0, bn0, sd1, bn1, sd2 = t0[0], t0[1], t0[3], t0[4], t1[1]
                    for owner, attr in ((sd0, 'weight'), (bn0, 'weight'), (bn0, 'bias'), (sd1, 'weight'),
                                        (sd2, 'weight'), (sd2, 'bias')):
                        t = owner._parameters[attr]
                        M['point'].add(owner, attr, True, off, t.shape)
                        off += t.numel()
                    boff = (j * L + l) * 8 * Fd + (0 if X == 'mu' else 4 * Fd)
                    for owner, attr in ((bn0, 'running_mean'), (bn0, 'running_var'), (bn1, 'running_mean'),
                                        (bn1, 'running_var')):
                        M['bn'].add(owner, attr, False, boff, (Fd,))
                        boff += Fd
                    for owner in (bn0, bn1):
                        M['nbt'].add(owner, 'num_batches_tracked', False, n, ())
                        n += 1
                    for which in ('w', 'b'):
                        seq = getattr(m, 'T_%s_0_cond_%s' % (X, which))
                        lin0, bn, lin1 = seq[0], seq[1], seq[3]
                        M['c_w0'].add(lin0, 'weight', True, c * Fd * G, (Fd, G))
                        M['c_bnw'].add(bn, 'weight', True, c * Fd, (Fd,))
                        M['c_bnb'].add(bn, 'bias', True, c * Fd, (Fd,))
                        M['c_rm'].add(bn, 'running_mean', False, c * Fd, (Fd,))
                        M['c_rv'].add(bn, 'running_var', False, c * Fd, (Fd,))
                        M['c_w1'].add(lin1, 'weight', True, c * Fd * Fd, (Fd, Fd))
                        M['c_b1'].add(lin1, 'bias', True, c * Fd, (Fd,))
                        M['nbt'].add(bn, 'num_batches_tracked', False, n, ())
                        n += 1
                        c += 1
        assert c == C and n == K * L * 8

    def _is_flat(self):
        for ms in self.masters.values():
            if ms.tensor is None:
                return False
            base = ms.tensor.data_ptr()
            esz = ms.tensor.element_size()
            for m in (ms.members[0], ms.members[-1]):
                if ms.current(m).data_ptr() != base + m[3] * esz:
                    return False
        return True

    @torch.no_grad()
    def flatten(self):
        """(Re)build the masters from the current module tensors and re-point the modules at them.
        Needed once, and again after `.to()/.cuda()` (which re-allocates every tensor separately)."""
        dev = self.layers[0][0].point[0].device
        for name, ms in self.masters.items():
            first = ms.current(ms.members[0])
            if first.dtype != ms.dtype:
                raise nat.GwtfError('the flow stack holds fp32 parameters (got %s)' % first.dtype)
            flat = torch.zeros(ms.shape, device=dev, dtype=ms.dtype)
            fv = flat.view(-1)
            for m in ms.members:
                owner, attr, is_param, off, shape = m
                cur = ms.current(m)
                numel = cur.numel()
                view = fv[off:off + numel].view(shape)
                view.copy_(cur.detach().to(dev))
                if is_param:
                    cur.data = view
                else:
                    owner._buffers[attr] = view
            if name in self.grad_masters:
                ms.params = [ms.current(m) for m in ms.members]
                for prm in ms.params:              # lets the optimizer update the whole master in one kernel
                    prm._gwtf_master = ms
                flat.requires_grad_(True)
                flat.register_hook(self._make_reduce_hook())
                flat.register_post_accumulate_grad_hook(self._make_attach_hook(ms))
            ms.tensor = flat
            ms.grad_views = None
            ms.grad_buf = None

    def _make_reduce_hook(self):
        def hook(grad):
            if self.sync_gradients and _world() > 1:
                grad = grad.contiguous()
                dist.all_reduce(grad)
                grad = grad / _world()
            return grad
        return hook

    def _make_attach_hook(self, ms):
        def hook(master):
            self._attach_grads(ms)
        return hook

    def _attach_grads(self, ms):
        """Give every module Parameter of a master its `.grad` as a view of the master's gradient.  The
        gradient lives in a persistent buffer (autograd hands over a fresh tensor after every
        `zero_grad(set_to_none=True)`; it is copied in), so the ~4,000 views are built once and a step only
        re-assigns them when the caller dropped the parameters' .grad."""
        g = ms.tensor.grad
        if g is None:
            return
        buf = ms.grad_buf
        if buf is None or buf.shape != g.shape or buf.device != g.device or buf.dtype != g.dtype:
            buf = ms.grad_buf = torch.empty_like(g, memory_format=torch.contiguous_format)
            gv = buf.view(-1)
            ms.grad_views = [gv[m[3]:m[3] + _numel(m[4])].view(m[4]) for m in ms.members]
        if g.data_ptr() != buf.data_ptr():
            buf.copy_(g)
            ms.tensor.grad = buf
        sentinel = ms.params[0]
        if sentinel.grad is None or sentinel.grad.data_ptr() != buf.data_ptr() + ms.members[0][3] * buf.element_size():
            for prm, v in zip(ms.params, ms.grad_views):
                prm.grad = v

    def _modules_replaced(self):
        """True when a BatchNorm of the stack was swapped for another module object after this stack was
        built (torch.nn.SyncBatchNorm.convert_sync_batchnorm, train_ae.py:152, creates new modules that
        share the old tensors)."""
        for comp in (self._components[0], self._components[-1]):
            for m in (comp[0], comp[-1]):
                ref = self.layers[self._components.index(comp)][comp.index(m)]
                if m.T_mu_0[1] is not ref.bn_point[0][0] or m.T_logvar_0_cond_b[1] is not ref.cond[3][1]:
                    return True
        return False

    def prepare(self):
        """Called at the top of every pass: keep the flat storage valid and honour
        `optimizer.zero_grad(set_to_none=True)` (parameters lost their .grad -> clear the masters)."""
        if self._modules_replaced():
            self.layers = [[_LayerRef(m) for m in comp] for comp in self._components]
            if self.flat:
                self._build_layout()
        if not self.flat:
            return
        if not self._is_flat():
            self.flatten()
        for name in self.grad_masters:
            ms = self.masters[name]
            if ms.tensor.grad is not None and ms.current(ms.members[0]).grad is None:
                ms.tensor.grad = None      # (the persistent buffer and its views are kept)

    def global_points(self, B, N, dev):
        """Number of points the SyncBN statistics run over (sum of B*N over ranks).  The all-reduced
        count needs a host read, so it is cached per local shape (ranks keep their batch sizes)."""
        key = (B, N, _world())
        if key not in self._n_total_cache:
            cnt = torch.tensor([float(B * N)], device=dev, dtype=torch.float64)
            dist.all_reduce(cnt)
            self._n_total_cache[key] = float(cnt.item())
        return self._n_total_cache[key]

    def pack_params(self):
        if self.flat:
            return self.masters['point'].tensor
        pieces = []
        for j in range(self.K):
            for l in range(self.L):
                ref = self.layers[j][l]
                n = sum(t.numel() for t in ref.point)
                pieces += [t.reshape(-1) for t in ref.point]
                if n < self.rec_stride:
                    pieces.append(ref.point[0].new_zeros(self.rec_stride - n))
        return torch.cat(pieces).view(self.K, self.L, self.rec_stride)

    def pack_bn(self):
        if self.flat:
            return self.masters['bn'].tensor
        pieces = []
        for j in range(self.K):
            for l in range(self.L):
                for bn0, bn1 in self.layers[j][l].bn_point:
                    pieces += [bn0.running_mean, bn0.running_var, bn1.running_mean, bn1.running_var]
        return torch.cat(pieces).view(self.K, self.L, 8 * self.F)

    def _cond_tensors(self):
        """name -> stacked tensor of the C FiLM nets (masters, or gathered per call)."""
        if self.flat:
            return {k: self.masters[k].tensor for k in ('c_w0', 'c_bnw', 'c_bnb', 'c_rm', 'c_rv', 'c_w1', 'c_b1')}
        conds = [c for j in range(self.K) for l in range(self.L) for c in self.layers[j][l].cond]
        return {'c_w0': torch.stack([c[0].weight for c in conds]), 'c_bnw': torch.stack([c[1].weight for c in conds]),
                'c_bnb': torch.stack([c[1].bias for c in conds]),
                'c_rm': torch.stack([c[1].running_mean for c in conds]),
                'c_rv': torch.stack([c[1].running_var for c in conds]),
                'c_w1': torch.stack([c[2].weight for c in conds]), 'c_b1': torch.stack([c[2].bias for c in conds])}

    @torch.no_grad()
    def update_point_bn(self, bstat, n_total):
        """nn.BatchNorm1d running-stat update from the batch statistics the kernels used.
        bstat (L,K,2,4,F): mean0 | var0 (biased) | mean1 | var1 (biased)."""
        unbias = float(n_total) / max(float(n_total) - 1.0, 1.0)
        K, L, Fd = self.K, self.L, self.F
        new = bstat.permute(1, 0, 2, 3, 4).clone()
        new[:, :, :, 1::2, :] *= unbias          # (no host tensor: an H2D copy here would sync the stream)
        if not self.flat:
            rs, ns, nbt = [], [], []
            for j in range(K):
                for l in range(L):
                    for net, (bn0, bn1) in enumerate(self.layers[j][l].bn_point):
                        rs += [bn0.running_mean, bn0.running_var, bn1.running_mean, bn1.running_var]
                        ns += list(new[j, l, net].unbind(0))
                        nbt += [bn0.num_batches_tracked, bn1.num_batches_tracked]
            torch._foreach_lerp_(rs, ns, BN_MOMENTUM)
            torch._foreach_add_(nbt, 1)
            return
        self.masters['bn'].tensor.view(K, L, 2, 4, Fd).lerp_(new, BN_MOMENTUM)
        # the first K*L*4 counters of each layer block belong to the point-wise BatchNorms
        nbt = self.masters['nbt'].tensor.view(K * L, 2, 4)
        nbt[:, :, :2] += 1

    # ------------------------------------------------------------------ FiLM nets
    def film(self, g, training, sync, update_stats=True):
        """(B,K,L,2,2,F): [...,0,:] = eps + exp(cond_w(g)), [...,1,:] = cond_b(g)."""
        K, L, Fd, C = self.K, self.L, self.F, self.C
        T = self._cond_tensors()
        H = F.linear(g, T['c_w0'].reshape(C * Fd, self.G))                                  # (B, C*F)
        bw, bb = T['c_bnw'].reshape(-1), T['c_bnb'].reshape(-1)
        rm, rv = T['c_rm'].reshape(-1), T['c_rv'].reshape(-1)
        if training:
            Hn, mean, var_unb = _batch_norm_train(H, bw, bb, sync)
            with torch.no_grad():
                if not update_stats:
                    pass
                elif self.flat:
                    rm.lerp_(mean, BN_MOMENTUM)
                    rv.lerp_(var_unb, BN_MOMENTUM)
                    self.masters['nbt'].tensor.view(K * L, 2, 4)[:, :, 2:] += 1
                else:
                    conds = [c for j in range(K) for l in range(L) for c in self.layers[j][l].cond]
                    torch._foreach_lerp_([c[1].running_mean for c in conds], list(mean.view(C, Fd).unbind(0)),
                                         BN_MOMENTUM)
                    torch._foreach_lerp_([c[1].running_var for c in conds], list(var_unb.view(C, Fd).unbind(0)),
                                         BN_MOMENTUM)
                    torch._foreach_add_([c[1].num_batches_tracked for c in conds], 1)
        else:
            Hn = (H - rm) * torch.rsqrt(rv + BN_EPS) * bw + bb
        A = Hn * torch.sigmoid(Hn)
        O = torch.baddbmm(T['c_b1'].unsqueeze(1), A.view(-1, C, Fd).transpose(0, 1),
                          T['c_w1'].transpose(1, 2))                                         # (C,B,F)
        O = O.view(K, L, 2, 2, -1, Fd).permute(4, 0, 1, 2, 3, 5)                            # (B,K,L,net,which,F)
        eps = self.layers[0][0].module.eps
        s = eps + torch.exp(O[..., 0, :])
        return torch.stack([s, O[..., 1, :]], dim=4).contiguous()

    # ------------------------------------------------------------------ kernels
    def nll_pass(self, p, g, training):
        """-> z (K,B,3,N) base-space samples, ssum (K,B,3,N) per-dim sums of logvar."""
        self.prepare()
        sync = training and _world() > 1
        film = self.film(g, training, sync)
        z, ssum, bstat, n_total = _StackNLLPass.apply(p.contiguous(), self.pack_params(), film, self.pack_bn(), self,
                                                      training, sync)
        if training:
            self.update_point_bn(bstat, n_total)
        return z, ssum

    # ------------------------------------------------------------------ options / diagnostics
    def fwd_engine(self):
        return int(nat.lib().gwtf_resolved_engine(ctypes.byref(self.desc), 0))

    def bwd_engine(self):
        return int(nat.lib().gwtf_resolved_engine(ctypes.byref(self.desc), 1))

    def nonfinite_counter(self, dev):
        """int32 device counter the NLL kernels bump for every non-finite per-point NLL they write."""
        if self._nonfinite is None or self._nonfinite.device != dev:
            self._nonfinite = torch.zeros(1, dtype=torch.int32, device=dev)
            self.desc.nonfinite = self._nonfinite.data_ptr()
        return self._nonfinite

    def take_nonfinite(self):
        """Number of non-finite per-point NLLs since the last call (one 4-byte read-back); the reference's
        loop stops without an optimizer step when the loss is NaN (training.py:43-46)."""
        if self._nonfinite is None:
            return 0
        n = int(self._nonfinite.item())
        if n:
            self._nonfinite.zero_()
        return n

    def _scratch(self, key, nbytes, dev):
        """Stream-ordered scratch reused across calls of the same shape (never handed to autograd)."""
        buf = self._keep_pool.get(key)
        if buf is None or buf.numel() < nbytes or buf.device != dev:
            buf = self._keep_pool[key] = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        return buf

    @torch.no_grad()
    def nll_eval_fused(self, p, g, base, logw, want_logp=False):
        """Fused no-grad eval forward -> nll (B,N) [, logp (B,N,K)]."""
        self.prepare()
        film = self.film(g, False, False)
        params = self.pack_params()
        bnbuf = self.pack_bn()
        p = p.contiguous()
        B, _, N = p.shape
        nll = torch.empty(B, N, device=p.device)
        logp = torch.empty(B, N, self.K, device=p.device) if want_logp else None
        self.nonfinite_counter(p.device)
        desc = ctypes.byref(self.desc)
        lib = nat.lib()
        if self.fwd_engine() != nat.ENGINE_FMA:
            # per-layer tensor-core kernels (L launches, two ping-pong slots): faster than the single-launch
            # FMA kernel at every size measured (64 x 2048: 2.2 vs 3.3 ms; 4 x 2048: 0.5 vs 3.3 ms)
            need = int(lib.gwtf_eval_layers_workspace_bytes(desc, B, N))
            ws = self._scratch('eval_ws', need, p.device)
            nat.check(lib.gwtf_nll_fwd_eval_layers(desc, nat.ptr(params), nat.ptr(bnbuf), nat.ptr(film), nat.ptr(p),
                                                   nat.ptr(base.contiguous()), nat.ptr(logw.contiguous()), nat.ptr(ws),
                                                   need, B, N, nat.ptr(nll), nat.ptr(logp), _stream_ptr()),
                      'gwtf_nll_fwd_eval_layers')
            return (nll, logp) if want_logp else nll
        nat.check(lib.gwtf_nll_fwd_eval(desc, nat.ptr(params), nat.ptr(bnbuf), nat.ptr(film),
                                        nat.ptr(p), nat.ptr(base.contiguous()), nat.ptr(logw.contiguous()),
                                        B, N, nat.ptr(nll), nat.ptr(logp), None, None, _stream_ptr()),
                  'gwtf_nll_fwd_eval')
        return (nll, logp) if want_logp else nll


def _numel(shape):
    n = 1
    for d in shape:
        n *= d
    return n


def _batch_norm_train(H, weight, bias, sync):
    """Train-mode BatchNorm over dim 0 of (B,C); returns normalised output and the (detached)
    batch mean / unbiased variance for the running-stat update.  With `sync` the statistics are
    shared by all ranks (SyncBatchNorm semantics, train_ae.py:152)."""
    if not sync:
        mean = H.mean(0)
        var = H.var(0, unbiased=False)
        n = H.shape[0]
        out = (H - mean) * torch.rsqrt(var + BN_EPS) * weight + bias
        return out, mean.detach(), (var * (n / max(n - 1, 1))).detach()
    return _SyncBN.apply(H, weight, bias)


class _SyncBN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, H, weight, bias):
        # sums in fp64: E[x^2] - mean^2 of fp32 sums loses the variance of channels whose mean dominates
        Hd = H.double()
        stats = torch.cat([Hd.sum(0), (Hd * Hd).sum(0), Hd.new_full((1,), float(H.shape[0]))])
        dist.all_reduce(stats)
        C = H.shape[1]
        nd = stats[-1]
        mean_d = stats[:C] / nd
        var_d = (stats[C:2 * C] / nd - mean_d * mean_d).clamp_min(0)
        mean, var, n = mean_d.to(H.dtype), var_d.to(H.dtype), nd.to(H.dtype)
        istd = torch.rsqrt(var + BN_EPS)
        xhat = (H - mean) * istd
        ctx.save_for_backward(xhat, weight, istd, n)
        ctx.mark_non_differentiable(mean, var)
        return xhat * weight + bias, mean, (var_d * (nd / (nd - 1).clamp_min(1))).to(H.dtype)

    @staticmethod
    def backward(ctx, dy, _dm, _dv):
        xhat, weight, istd, n = ctx.saved_tensors
        dw = (dy * xhat).sum(0)
        db = dy.sum(0)
        C = dy.shape[1]
        sums = torch.cat([db, dw])
        dist.all_reduce(sums)
        dxhat = dy * weight
        mean_d = sums[:C] * weight / n
        mean_dx = sums[C:] * weight / n
        dH = (dxhat - mean_d - xhat * mean_dx) * istd
        return dH, dw, db


class _StackNLLPass(torch.autograd.Function):
    """points (B,3,N), params (K,L,RS), film (B,K,L,2,2,F) -> z, ssum  (reference mode='inverse',
    decoders.py:72-77 over all K components)."""

    @staticmethod
    def forward(ctx, p, params, film, bnbuf, stack, training, sync):
        lib = nat.lib()
        K, L, Fd = stack.K, stack.L, stack.F
        B, _, N = p.shape
        dev = p.device
        desc = ctypes.byref(stack.desc)
        st = _stream_ptr()
        ubuf = torch.empty(L, K, B, 3, N, device=dev)
        ssum = torch.empty(K, B, 3, N, device=dev)      # (the drivers zero what they accumulate into)
        ld = torch.empty(K, B, N, device=dev)
        mom = sum1 = bstat = None
        n_total = float(B * N)
        # activations kept for backward (layout private to the engine, sized by the library)
        ybuf = None
        keep = bool(stack.keep_activations) and any(ctx.needs_input_grad) and \
            int(lib.gwtf_keep_floats(desc, B, N)) > 0
        if keep:
            # multi-GB scratch: recycled through a per-stack pool (returned by backward) so that the
            # caching allocator never splits or re-mallocs it between steps
            need = int(lib.gwtf_keep_floats(desc, B, N))
            pool = stack._keep_pool.setdefault((need, dev), [])
            if pool:
                ybuf = pool.pop()
            else:
                free, _ = torch.cuda.mem_get_info(dev)
                if 4 * need < 0.5 * free + torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev):
                    ybuf = torch.empty(need, device=dev)
        if training:
            mom = torch.empty(L, K, nat.MOM_STRIDE, device=dev, dtype=torch.float64)
            sum1 = torch.empty(L, K, 2, 2, Fd, device=dev, dtype=torch.float64)
            bstat = torch.empty(L, K, 2, 4, Fd, device=dev)
        peer = peer_exchange(K * 8 * Fd, dev) if sync else None
        stack.desc.exchange = peer
        if not sync:
            nat.check(lib.gwtf_fwd_all(desc, int(training), nat.ptr(params), nat.ptr(bnbuf), nat.ptr(film), nat.ptr(p),
                                       None, None, nat.ptr(ubuf), nat.ptr(ld), nat.ptr(ssum), nat.ptr(ybuf), nat.ptr(mom),
                                       nat.ptr(sum1), nat.ptr(bstat), B, N, None, None, st), 'gwtf_fwd_all')
        elif peer:
            # one call: the statistic sums cross ranks through NVLink peer memory between the kernel phases
            n_total = stack.global_points(B, N, dev)
            nat.check(lib.gwtf_fwd_all_ranks(desc, 1, nat.ptr(params), nat.ptr(bnbuf), nat.ptr(film), nat.ptr(p),
                                             None, None, nat.ptr(ubuf), nat.ptr(ld), nat.ptr(ssum), nat.ptr(ybuf),
                                             nat.ptr(mom), nat.ptr(sum1), nat.ptr(bstat), B, N, None, None, n_total, st),
                      'gwtf_fwd_all_ranks')
        else:
            n_total = stack.global_points(B, N, dev)
            ssum.zero_()
            ld.zero_()
            mom.zero_()
            sum1.zero_()
            nat.check(lib.gwtf_fwd_moments(desc, nat.ptr(p), B, N, nat.ptr(mom[L - 1]), st), 'gwtf_fwd_moments')
            for l in range(L - 1, -1, -1):
                dist.all_reduce(mom[l])
                for phase in (0, 1):
                    nat.check(lib.gwtf_fwd_layer(desc, l, phase, 1, nat.ptr(params), nat.ptr(bnbuf), nat.ptr(film),
                                                 nat.ptr(p), nat.ptr(ubuf), nat.ptr(ld), nat.ptr(ssum), nat.ptr(ybuf),
                                                 nat.ptr(mom), nat.ptr(sum1), B, N, n_total, st), 'gwtf_fwd_layer')
                    if phase == 0:
                        dist.all_reduce(sum1[l])
            nat.check(lib.gwtf_fwd_bstat(desc, nat.ptr(params), nat.ptr(mom), nat.ptr(sum1), n_total, nat.ptr(bstat),
                                         st), 'gwtf_fwd_bstat')
        ctx.stack = stack
        ctx.training = training
        ctx.sync = sync
        ctx.peer = peer is not None
        ctx.n_total = n_total
        # running stats are updated in place right after a train-mode forward (and are not read by
        # the train-mode backward), so they must not go through save_for_backward's version check
        ctx.bnbuf = bnbuf
        ctx.ybuf = ybuf
        ctx.save_for_backward(p, params, film, ubuf, mom, sum1)
        z = ubuf[0]
        ctx.mark_non_differentiable(*([bstat] if bstat is not None else []))
        return z, ssum, bstat, n_total

    @staticmethod
    def backward(ctx, dz, dssum, _dbstat, _dn):
        lib = nat.lib()
        stack = ctx.stack
        p, params, film, ubuf, mom, sum1 = ctx.saved_tensors
        bnbuf = ctx.bnbuf
        ybuf = ctx.ybuf
        K, L, Fd = stack.K, stack.L, stack.F
        B, _, N = p.shape
        dev = p.device
        desc = ctypes.byref(stack.desc)
        st = _stream_ptr()
        gbuf = dz.contiguous().clone() if dz is not None else torch.zeros(K, B, 3, N, device=dev)
        gs = dssum.contiguous() if dssum is not None else torch.zeros(K, B, 3, N, device=dev)
        dobuf = torch.empty(K, B, 6, N, device=dev)
        acc = torch.zeros(params.numel() + film.numel() + p.numel(), device=dev)      # one memset for the three
        dparams = acc[:params.numel()].view_as(params)
        dfilm = acc[params.numel():params.numel() + film.numel()].view_as(film)
        dpoints = acc[params.numel() + film.numel():].view_as(p)
        bsum = torch.zeros(L, K, 2, 4, Fd, device=dev, dtype=torch.float64)
        train = int(ctx.training)
        if not ctx.sync:
            nat.check(lib.gwtf_bwd_all(desc, train, nat.ptr(params), nat.ptr(bnbuf), nat.ptr(film), nat.ptr(p),
                                       None, None, nat.ptr(ubuf), nat.ptr(ybuf), None, nat.ptr(mom), nat.ptr(sum1), None, None,
                                       nat.ptr(bsum), nat.ptr(gbuf), nat.ptr(gs), nat.ptr(dobuf), nat.ptr(dparams),
                                       nat.ptr(dfilm), None, None, nat.ptr(dpoints), B, N, st), 'gwtf_bwd_all')
        elif ctx.peer:
            nat.check(lib.gwtf_bwd_all_ranks(desc, train, nat.ptr(params), nat.ptr(bnbuf), nat.ptr(film), nat.ptr(p),
                                             None, None, nat.ptr(ubuf), nat.ptr(ybuf), None, nat.ptr(mom), nat.ptr(sum1),
                                             None, None, nat.ptr(bsum), nat.ptr(gbuf), nat.ptr(gs), nat.ptr(dobuf),
                                             nat.ptr(dparams), nat.ptr(dfilm), None, None, nat.ptr(dpoints), B, N,
                                             ctx.n_total, st), 'gwtf_bwd_all_ranks')
        else:
            for l in range(L):
                for phase in (0, 1):
                    nat.check(lib.gwtf_bwd_layer(desc, l, phase, train, nat.ptr(params), nat.ptr(bnbuf),
                                                 nat.ptr(film), nat.ptr(p), nat.ptr(ubuf), nat.ptr(ybuf), nat.ptr(mom),
                                                 nat.ptr(sum1), nat.ptr(bsum), nat.ptr(gbuf), nat.ptr(gs),
                                                 nat.ptr(dobuf), nat.ptr(dparams), nat.ptr(dfilm), B, N,
                                                 ctx.n_total, st), 'gwtf_bwd_layer')
                    dist.all_reduce(bsum[l])
            nat.check(lib.gwtf_bwd_finish(desc, train, nat.ptr(params), nat.ptr(bnbuf), nat.ptr(mom), nat.ptr(bsum),
                                          nat.ptr(gbuf), nat.ptr(p), nat.ptr(dparams), nat.ptr(dpoints), B, N,
                                          ctx.n_total, st), 'gwtf_bwd_finish')
        if ybuf is not None:            # stream-ordered reuse: the next forward runs after these launches
            stack._keep_pool.setdefault((ybuf.numel(), dev), []).append(ybuf)
            ctx.ybuf = None
        return dpoints, dparams, dfilm, None, None, None, None


class _MixtureHead(torch.autograd.Function):
    """z, ssum, base (B,2,3), logw (B,K) -> per-point mixture NLL (B,N)  (losses.py:112-128)."""

    @staticmethod
    def forward(ctx, z, ssum, base, logw, stack):
        lib = nat.lib()
        K, B, _, N = z.shape
        dev = z.device
        z = z.contiguous()
        ld = ssum.sum(2)
        base = base.contiguous()
        logw = logw.contiguous()
        nll = torch.empty(B, N, device=dev)
        stack.nonfinite_counter(dev)
        nat.check(lib.gwtf_nll_from_state(ctypes.byref(stack.desc), nat.ptr(z), nat.ptr(ld), nat.ptr(base),
                                          nat.ptr(logw), B, N, nat.ptr(nll), None, _stream_ptr()),
                  'gwtf_nll_from_state')
        ctx.stack = stack
        ctx.save_for_backward(z, ld, base, logw, nll)
        return nll

    @staticmethod
    def backward(ctx, dnll):
        lib = nat.lib()
        z, ld, base, logw, nll = ctx.saved_tensors
        K, B, _, N = z.shape
        dev = z.device
        gbuf = torch.empty(K, B, 3, N, device=dev)
        gs = torch.empty(K, B, 3, N, device=dev)
        dbase = torch.zeros_like(base)
        dlogw = torch.zeros_like(logw)
        nat.check(lib.gwtf_bwd_seed(ctypes.byref(ctx.stack.desc), nat.ptr(z), nat.ptr(ld), nat.ptr(base),
                                    nat.ptr(logw), nat.ptr(nll), nat.ptr(dnll.contiguous()), B, N, nat.ptr(gbuf),
                                    nat.ptr(gs), nat.ptr(dbase), nat.ptr(dlogw), _stream_ptr()), 'gwtf_bwd_seed')
        return gbuf, gs, dbase, dlogw, None


# ---------------------------------------------------------------------------------------------
# public entry points used by the drop-in modules
# ---------------------------------------------------------------------------------------------
def mixture_nll(stack, p, g, mu_base, lv_base, logits, training, want_nll=True):
    """All K inverse stacks + (optionally) the in-kernel mixture NLL.

    -> (z (K,B,3,N), ssum (K,B,3,N), nll (B,N) or None).  With gradients disabled and eval-mode
    BatchNorm the single fused kernel is used and z / ssum are not materialised.
    """
    nat.lib()
    if not p.is_cuda:
        raise nat.GwtfError('the flow stack runs on CUDA tensors only (got %s)' % p.device)
    base = torch.stack([mu_base, lv_base], dim=1)                                   # (B,2,3)
    logw = logits - torch.logsumexp(logits, dim=-1, keepdim=True)                   # losses.py:101-104
    needs_grad = torch.is_grad_enabled()
    if want_nll and not training and not needs_grad:
        return None, None, stack.nll_eval_fused(p, g, base, logw)
    z, ssum = stack.nll_pass(p, g, training)
    nll = _MixtureHead.apply(z, ssum, base, logw, stack) if want_nll else None
    return z, ssum, nll


def mixture_cdf(logits):
    """(B,K) logits on the device -> (B,K) inclusive CDF of softmax(logits) the way np.random.choice builds it
    (flow_mixture.py:149-153): fp32 probabilities, float64 cumsum, normalised; stored fp32 with the last entry
    pinned to 1.  Computed by a kernel (no host round trip; the reference syncs to numpy here, :149)."""
    logits = logits.detach().float().contiguous()
    cdf = torch.empty_like(logits)
    nat.check(nat.lib().gwtf_mixture_cdf(nat.ptr(logits), logits.shape[0], logits.shape[1], nat.ptr(cdf), _stream_ptr()),
              'gwtf_mixture_cdf')
    return cdf


@torch.no_grad()
def sample_mixture(stack, g, mu_base, lv_base, logits, n_points, seed, stream_id=0, idx=None, eps=None,
                   want_z=False):
    """Eval-mode sampling of n_points per shape -> samples (B,3,N), labels (B,N) int32 in 1..K,
    z (B,3,N) or None.  `idx` (B,N) int32 / `eps` (B,3,N) replace the in-kernel Philox draws.
    Nothing is read back to the host: the call only enqueues kernels on the current stream."""
    lib = nat.lib()
    if not g.is_cuda:
        raise nat.GwtfError('the flow stack runs on CUDA tensors only (got %s)' % g.device)
    B = g.shape[0]
    dev = g.device
    stack.prepare()
    film = stack.film(g, False, False)
    params = stack.pack_params()
    bnbuf = stack.pack_bn()
    base = torch.stack([mu_base, lv_base], dim=1).contiguous()
    cdf = mixture_cdf(logits)
    samples = torch.empty(B, 3, n_points, device=dev)
    labels = torch.empty(B, n_points, device=dev, dtype=torch.int32)
    z = torch.empty(B, 3, n_points, device=dev) if want_z else None
    if idx is not None:
        idx = idx.to(device=dev, dtype=torch.int32).contiguous()
    if eps is not None:
        eps = eps.to(device=dev, dtype=torch.float32).contiguous()
    desc = ctypes.byref(stack.desc)
    seed64, sid = ctypes.c_uint64(seed & (2 ** 64 - 1)), ctypes.c_uint32(stream_id & 0xFFFFFFFF)
    if stack.fwd_engine() == nat.ENGINE_TC_FWD:
        # tcgen05 forward: regroup the points by component inside each shape's row and run the per-layer kernels
        # in direct mode; every buffer is sized from upper bounds, so there is no count read-back
        need = int(lib.gwtf_sample_workspace_bytes(desc, B, n_points))
        ws = stack._scratch('sample_ws', need, dev)
        nat.check(lib.gwtf_sample_layers(desc, nat.ptr(params), nat.ptr(bnbuf), nat.ptr(film), nat.ptr(base),
                                         nat.ptr(cdf), B, n_points, seed64, sid, nat.ptr(idx), nat.ptr(eps), nat.ptr(ws),
                                         need, nat.ptr(samples), nat.ptr(labels), nat.ptr(z), _stream_ptr()),
                  'gwtf_sample_layers')
        return samples, labels, z
    nat.check(lib.gwtf_sample(desc, nat.ptr(params), nat.ptr(bnbuf), nat.ptr(film), nat.ptr(base),
                              nat.ptr(cdf), B, n_points, seed64, sid, nat.ptr(idx), nat.ptr(eps), nat.ptr(samples),
                              nat.ptr(labels), nat.ptr(z), _stream_ptr()), 'gwtf_sample')
    return samples, labels, z


def run_module_stack(stack, p, g, mode, training):
    """Per-module list API (flows.py:117,160; decoders.py:79) on a K=1 stack: every layer's
    (p_out, mu, logvar), indexed by layer in DIRECT order.

    Gradients.  mode='inverse' with autograd enabled is differentiable in what the reference's losses consume
    (losses.py:12-20,112-122): ps[0], the base-space sample, carries the graph, and so does the SUM of the
    returned logvars (logvars[0] is returned as `S - sum(logvars[1:])` with S the differentiable per-dim
    log-det sum of the kernels; the other entries are values only).  Per-layer intermediates (ps[1:], mus) are
    values only.  mode='direct' is the sampling direction and never carries a graph; asking for it in training
    mode with autograd enabled raises instead of silently dropping gradients."""
    if mode not in ('direct', 'inverse'):
        raise ValueError(mode)
    if torch.is_grad_enabled():
        if mode == 'direct' and training:
            raise nat.GwtfError("mode='direct' (sampling) is not differentiable here: call it under torch.no_grad() "
                                "or in eval mode; training goes through mode='inverse' / Flow_Mixture_Model.decode")
        if mode == 'inverse':
            z, ssum = stack.nll_pass(p.contiguous().float(), g, training)          # (1,B,3,N) each, with graph
            ps, mus, lvs = _run_module_stack(stack, p, g, mode, training, update_stats=False)
            ps = [z[0]] + ps[1:]
            if len(lvs) > 1:
                rest = torch.stack(lvs[1:]).sum(0)
                lvs = [ssum[0] - rest] + lvs[1:]
            else:
                lvs = [ssum[0]]
            return ps, mus, lvs
    return _run_module_stack(stack, p, g, mode, training, update_stats=True)


@torch.no_grad()
def _run_module_stack(stack, p, g, mode, training, update_stats):
    """The list API proper (values only)."""
    lib = nat.lib()
    if not p.is_cuda:
        raise nat.GwtfError('the flow stack runs on CUDA tensors only (got %s)' % p.device)
    assert stack.K == 1
    L, Fd = stack.L, stack.F
    p = p.contiguous().float()
    B, _, N = p.shape
    dev = p.device
    sync = training and _world() > 1
    stack.prepare()
    film = stack.film(g, training, sync, update_stats=update_stats)
    params = stack.pack_params()
    bnbuf = stack.pack_bn()
    desc = ctypes.byref(stack.desc)
    st = _stream_ptr()
    trio = torch.empty(L, 3, B, 3, N, device=dev)
    order = list(range(L)) if mode == 'direct' else list(range(L - 1, -1, -1))
    n_total = float(B * N)
    mom = sum1 = None
    if training:
        mom = torch.zeros(L, 1, nat.MOM_STRIDE, device=dev, dtype=torch.float64)
        sum1 = torch.zeros(L, 1, 2, 2, Fd, device=dev, dtype=torch.float64)
        if sync:
            n_total = stack.global_points(B, N, dev)
        nat.check(lib.gwtf_fwd_moments(desc, nat.ptr(p), B, N, nat.ptr(mom[order[0]]), st), 'gwtf_fwd_moments')
    cur = p
    for i, l in enumerate(order):
        nxt = order[i + 1] if i + 1 < L else None
        phases = (0, 1) if training else (1,)
        if sync:
            dist.all_reduce(mom[l])
        for phase in phases:
            xout = trio[l, 0]
            nat.check(lib.gwtf_fwd_layer_ex(desc, l, phase, int(training), int(mode == 'direct'), nat.ptr(params),
                                            nat.ptr(bnbuf), nat.ptr(film), nat.ptr(cur), 1, nat.ptr(xout), None, None,
                                            nat.ptr(trio[l]), None, nat.ptr(mom[l]) if training else None,
                                            nat.ptr(mom[nxt]) if (training and nxt is not None) else None,
                                            nat.ptr(sum1[l]) if training else None, B, N, n_total, st),
                      'gwtf_fwd_layer_ex')
            if sync and phase == 0:
                dist.all_reduce(sum1[l])
        cur = trio[l, 0]
    if training:
        bstat = torch.empty(L, 1, 2, 4, Fd, device=dev)
        nat.check(lib.gwtf_fwd_bstat(desc, nat.ptr(params), nat.ptr(mom), nat.ptr(sum1), n_total, nat.ptr(bstat), st),
                  'gwtf_fwd_bstat')
        if update_stats:
            stack.update_point_bn(bstat, n_total)
    ps = [trio[l, 0] for l in range(L)]
    mus = [trio[l, 1] for l in range(L)]
    lvs = [trio[l, 2] for l in range(L)]
    return ps, mus, lvs
