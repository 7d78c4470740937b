"""Optimizer + schedule.  Mirrors lib/networks/optimizers.py: Adam (AMSGrad, weight decay added to
the update un-scaled by lr, :69-72) and the cosine LRUpdater :79-97.  Stays PyTorch this round
(SURVEY.md §8 f2); the update is expressed with multi-tensor `torch._foreach_*` ops so one step is
a handful of launches instead of ~12 per parameter tensor."""
import math

import numpy as np
import torch
from torch.optim import Optimizer


class Adam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            buckets = {}
            for p in group['params']:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError('Adam does not support sparse gradients, please consider SparseAdam instead')
                st = self.state[p]
                if len(st) == 0:
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
