"""Optimizer + schedule.  Mirrors lib/networks/optimizers.py: Adam (AMSGrad, weight decay added to the
update un-scaled by lr, :69-72) and the cosine LRUpdater :79-97.

The decoder's parameters are views into a few flat master tensors (flowstack.FlowStack); for those the whole
step is ONE launch of the fused AMSGrad kernel per master (`gwtf_adam_step`, HBM-bound: 36 bytes per parameter)
with flat optimizer state whose per-parameter entries are views, so `state_dict()` keeps the reference's
per-parameter layout.  Everything else (encoders, prior flow: a few hundred small tensors) is updated with
multi-tensor `torch._foreach_*` ops -- a handful of launches instead of ~12 per tensor."""
import ctypes
import math

import numpy as np
import torch
from torch.optim import Optimizer


class Adam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad))
        self._flat = {}          # id(master) -> flat state of a FlowStack master

    # ------------------------------------------------------------------ flat masters
    @staticmethod
    def _master_ready(ms, plist):
        """All members present, storage still flat, gradients are the views of the master's gradient buffer."""
        if ms.tensor is None or ms.tensor.grad is None or ms.params is None or len(plist) != len(ms.params):
            return False
        if not ms.tensor.is_cuda or ms.tensor.dtype != torch.float32:
            return False
        first, off = ms.params[0], ms.members[0][3]
        base, gbase = ms.tensor.data_ptr(), ms.tensor.grad.data_ptr()
        return first.data_ptr() == base + 4 * off and first.grad is not None and first.grad.data_ptr() == gbase + 4 * off

    def _flat_state(self, ms, amsgrad):
        st = self._flat.get(id(ms))
        if st is None or st['exp_avg'].shape != ms.tensor.shape or st['exp_avg'].device != ms.tensor.device:
            st = {'step': 0, 'exp_avg': torch.zeros_like(ms.tensor, requires_grad=False),
                  'exp_avg_sq': torch.zeros_like(ms.tensor, requires_grad=False), 'master': ms}
            if amsgrad:
                st['max_exp_avg_sq'] = torch.zeros_like(ms.tensor, requires_grad=False)
            # adopt per-parameter state that already exists (a loaded checkpoint), then re-point it at the views
            names = [k for k in ('exp_avg', 'exp_avg_sq', 'max_exp_avg_sq') if k in st]
            flat = {k: st[k].view(-1) for k in names}
            for prm, m in zip(ms.params, ms.members):
                off, n = m[3], prm.numel()
                old = self.state.get(prm, {})
                if old:
                    st['step'] = max(st['step'], int(old.get('step', 0)))
                pst = self.state[prm]
                for k in names:
                    view = flat[k][off:off + n].view(prm.shape)
                    if k in old and torch.is_tensor(old[k]):
                        view.copy_(old[k])
                    pst[k] = view
                pst['step'] = st['step']
            self._flat[id(ms)] = st
        return st

    def _sync_steps(self):
        for st in self._flat.values():
            for prm in st['master'].params:
                self.state[prm]['step'] = st['step']

    def state_dict(self):
        self._sync_steps()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._flat = {}          # re-adopt the loaded per-parameter state at the next step

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            masters, rest = {}, []
            for p in group['params']:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError('Adam does not support sparse gradients, please consider SparseAdam instead')
                ms = getattr(p, '_gwtf_master', None)
                if ms is not None and p.is_cuda:
                    masters.setdefault(id(ms), (ms, []))[1].append(p)
                else:
                    rest.append(p)
            for ms, plist in masters.values():
                if not self._master_ready(ms, plist):
                    rest += plist
                    continue
                from .. import _native as nat
                st = self._flat_state(ms, group['amsgrad'])
                st['step'] += 1
                vmax = st.get('max_exp_avg_sq')
                nat.check(nat.lib().gwtf_adam_step(
                    nat.ptr(ms.tensor), nat.ptr(ms.tensor.grad), nat.ptr(st['exp_avg']), nat.ptr(st['exp_avg_sq']),
                    nat.ptr(vmax), ms.tensor.numel(), float(group['lr']), float(beta1), float(beta2), float(group['eps']),
                    float(group['weight_decay']), st['step'],
                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), 'gwtf_adam_step')
            buckets = {}
            for p in rest:
                st = self.state[p]
                if len(st) == 0 or 'exp_avg' not in st:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p)
                    st['exp_avg_sq'] = torch.zeros_like(p)
                    if group['amsgrad']:
                        st['max_exp_avg_sq'] = torch.zeros_like(p)
                st['step'] += 1
                buckets.setdefault((st['step'], p.device, p.dtype), []).append(p)
            for (step, _, _), ps in buckets.items():
                grads = [p.grad for p in ps]
                m = [self.state[p]['exp_avg'] for p in ps]
                v = [self.state[p]['exp_avg_sq'] for p in ps]
                torch._foreach_mul_(m, beta1)
                torch._foreach_add_(m, grads, alpha=1 - beta1)
                torch._foreach_mul_(v, beta2)
                torch._foreach_addcmul_(v, grads, grads, value=1 - beta2)
                if group['amsgrad']:
                    vmax = [self.state[p]['max_exp_avg_sq'] for p in ps]
                    torch._foreach_maximum_(vmax, v)
                    denom = torch._foreach_sqrt(vmax)
                else:
                    denom = torch._foreach_sqrt(v)
                bc1 = 1 - beta1 ** step
                bc2 = math.sqrt(1 - beta2 ** step)
                torch._foreach_div_(denom, bc2)
                torch._foreach_add_(denom, group['eps'])
                m_hat = torch._foreach_div(m, bc1)
                if group['weight_decay'] != 0:
                    # p -= wd * p + lr * m_hat / denom      (decay NOT scaled by lr, :69-72)
                    upd = torch._foreach_mul(ps, group['weight_decay'])
                    torch._foreach_addcdiv_(upd, m_hat, denom, value=group['lr'])
                    torch._foreach_sub_(ps, upd)
                else:
                    torch._foreach_addcdiv_(ps, m_hat, denom, value=-group['lr'])
        return loss


class LRUpdater(object):
    def __init__(self, epoch_length, **kwargs):
        self.epoch_length = epoch_length
        self.cycle_length = kwargs['cycle_length']
        self.min_lr = kwargs['min_lr']
        self.max_lr = kwargs['max_lr']
        self.beta1 = kwargs['beta1']
        self.min_beta2 = kwargs['min_beta2']
        self.max_beta2 = kwargs['max_beta2']

    def __call__(self, optimizer, epoch, iteration):
        rel_epoch = epoch % self.cycle_length
        t = (rel_epoch * self.epoch_length + iteration) / (self.cycle_length * self.epoch_length)
        cos = 0.5 * (1.0 + np.cos(np.pi * t))
        lr = self.min_lr + (self.max_lr - self.min_lr) * cos
        beta2 = self.min_beta2 + (self.max_beta2 - self.min_beta2) * cos
        for group in optimizer.param_groups:
            group['lr'] = lr
            group['betas'] = (self.beta1, beta2)
