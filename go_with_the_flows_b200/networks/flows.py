"""Flow modules.  Mirrors lib/networks/flows.py: CondRealNVPFlow3D :10-117, CondRealNVPFlow3DTriple
:120-160 (the hot path) and RealNVPFlow / RealNVPFlowCouple :163-243 (latent prior flow, PyTorch).

The conditional coupling layers only HOLD parameters (same names, shapes and init order as the
reference, so `load_state_dict(strict=True)` of reference checkpoints works); their arithmetic
runs in the sm_100a kernels of libgwtf.so through `FlowStack`.  `forward(p, g, mode)` keeps the
reference's list-returning contract (flows.py:117,160) for callers that use modules directly; the
training / evaluation loops go through `Flow_Mixture_Model.decode`, which hands all K stacks to
the fused kernels at once.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from .layers import SharedDot, Swish
from ..flowstack import FlowStack, run_module_stack


def _film_net(stem, g_features, f_features):
    return nn.Sequential(OrderedDict([
        (stem + '0', nn.Linear(g_features, f_features, bias=False)),
        (stem + '0_bn', nn.BatchNorm1d(f_features)),
        (stem + '0_swish', Swish()),
        (stem + '1', nn.Linear(f_features, f_features, bias=True)),
    ]))


def _small_init(lin, std):
    lin.weight.data.normal_(std=std)
    nn.init.constant_(lin.bias.data, 0.0)


class CondRealNVPFlow3D(nn.Module):
    """One FiLM-conditioned affine coupling layer on 3-D points (flows.py:10-117)."""

    def __init__(self, f_n_features, g_n_features, weight_std=0.01, warp_inds=[0],
                 centered_translation=False, eps=1e-6):
        super().__init__()
        self.f_n_features = f_n_features
        self.g_n_features = g_n_features
        self.weight_std = weight_std
        self.warp_inds = list(warp_inds)
        self.keep_inds = [d for d in (0, 1, 2) if d not in self.warp_inds]
        self.centered_translation = centered_translation      # stored, never used (flows.py:20)
        self.register_buffer('eps', torch.from_numpy(np.array([eps], dtype=np.float32)))
        for X in ('mu', 'logvar'):
            self._build_net(X)
        self._stack = None

    def _build_net(self, X):
        Fd, G = self.f_n_features, self.g_n_features
        k, w = len(self.keep_inds), len(self.warp_inds)
        setattr(self, 'T_%s_0' % X, nn.Sequential(OrderedDict([
            (X + '_sd0', SharedDot(k, Fd, 1)),
            (X + '_sd0_bn', nn.BatchNorm1d(Fd)),
            (X + '_sd0_relu', nn.ReLU(inplace=True)),
            (X + '_sd1', SharedDot(Fd, Fd, 1)),
            (X + '_sd1_bn', nn.BatchNorm1d(Fd, affine=False)),
        ])))
        cond_w = _film_net(X + '_sd1_film_w', G, Fd)
        setattr(self, 'T_%s_0_cond_w' % X, cond_w)
        cond_b = _film_net(X + '_sd1_film_b', G, Fd)
        setattr(self, 'T_%s_0_cond_b' % X, cond_b)
        head = nn.Sequential(OrderedDict([
            (X + '_sd1_relu', nn.ReLU(inplace=True)),
            (X + '_sd2', SharedDot(Fd, w, 1, bias=True)),
        ]))
        setattr(self, 'T_%s_1' % X, head)
        with torch.no_grad():
            _small_init(cond_w[-1], self.weight_std)
            _small_init(cond_b[-1], self.weight_std)
            _small_init(head[-1], self.weight_std)

    def forward(self, p, g, mode='direct'):
        if self._stack is None:
            self._stack = FlowStack([[self]], own_storage=False)
        ps, mus, lvs = run_module_stack(self._stack, p, g, mode, self.training)
        return ps[0], mus[0], lvs[0]


class CondRealNVPFlow3DTriple(nn.Module):
    """Three coupling layers; pattern 0 warps {0},{1},{2}, pattern 1 warps {0,1},{0,2},{1,2}
    (flows.py:129-148)."""

    _WARPS = {0: ([0], [1], [2]), 1: ([0, 1], [0, 2], [1, 2])}

    def __init__(self, f_n_features, g_n_features, weight_std=0.02, pattern=0, centered_translation=False):
        super().__init__()
        self.f_n_features = f_n_features
        self.g_n_features = g_n_features
        self.weight_std = weight_std
        self.pattern = pattern
        self.centered_translation = centered_translation
        if pattern in self._WARPS:
            for i, warp in enumerate(self._WARPS[pattern]):
                setattr(self, 'nvp%d' % (i + 1), CondRealNVPFlow3D(
                    f_n_features, g_n_features, weight_std=weight_std, warp_inds=list(warp),
                    centered_translation=centered_translation))
        self._stack = None

    def coupling_layers(self):
        return [self.nvp1, self.nvp2, self.nvp3]

    def forward(self, p, g, mode='direct'):
        if self._stack is None:
            self._stack = FlowStack([self.coupling_layers()], own_storage=False)
        # both modes return [p1, p2, p3] indexed by layer (flows.py:160)
        return run_module_stack(self._stack, p, g, mode, self.training)


# ---------------------------------------------------------------------------------------------
# latent prior flow (B x G tensors; negligible FLOPs -- stays PyTorch)
# ---------------------------------------------------------------------------------------------
class RealNVPFlow(nn.Module):
    def __init__(self, n_features, g_n_features, weight_std=0.01, warp_inds=[0], eps=1e-6):
        super().__init__()
        self.n_features = n_features
        self.g_n_features = g_n_features
        self.weight_std = weight_std
        self.warp_inds = list(warp_inds)
        warped = set(int(i) for i in self.warp_inds)
        self.keep_inds = [i for i in range(g_n_features) if i not in warped]
        self.register_buffer('eps', torch.from_numpy(np.array([eps], dtype=np.float32)))
        for X in ('mu', 'logvar'):
            net = nn.Sequential(OrderedDict([
                (X + '_mlp0', nn.Linear(len(self.keep_inds), n_features, bias=False)),
                (X + '_mlp0_bn', nn.BatchNorm1d(n_features)),
                (X + '_mlp0_swish', Swish()),
                (X + '_mlp1', nn.Linear(n_features, len(self.warp_inds), bias=True)),
            ]))
            with torch.no_grad():
                _small_init(net[-1], weight_std)
            setattr(self, 'T_%s_0' % X, net)

    def forward(self, g, mode='direct'):
        kept = g[:, self.keep_inds].contiguous()
        logvar = torch.zeros_like(g)
        mu = torch.zeros_like(g)
        logvar[:, self.warp_inds] = torch.log(self.eps + torch.exp(self.T_logvar_0(kept)))
        mu[:, self.warp_inds] = self.T_mu_0(kept)
        if mode == 'direct':
            g_out = torch.exp(0.5 * logvar) * g + mu
        elif mode == 'inverse':
            g_out = torch.exp(-0.5 * logvar) * (g - mu)
        else:
            raise ValueError(mode)
        return g_out, mu, logvar


class RealNVPFlowCouple(nn.Module):
    def __init__(self, n_features, g_n_features, weight_std=0.01, pattern=0):
        super().__init__()
        self.n_features = n_features
        self.g_n_features = g_n_features
        self.weight_std = weight_std
        self.pattern = pattern
        idx = np.arange(g_n_features)
        if pattern == 0:
            halves = (idx[::2], idx[1::2])
        elif pattern == 1:
            halves = (idx[:g_n_features // 2], idx[g_n_features // 2:])
        else:
            halves = ()
        for i, h in enumerate(halves):
            setattr(self, 'nvp%d' % (i + 1), RealNVPFlow(n_features, g_n_features, weight_std=weight_std,
                                                         warp_inds=list(h)))

    def forward(self, g, mode='direct'):
        if mode == 'direct':
            g1, mu1, lv1 = self.nvp1(g, mode=mode)
            g2, mu2, lv2 = self.nvp2(g1, mode=mode)
        elif mode == 'inverse':
            g2, mu2, lv2 = self.nvp2(g, mode=mode)
            g1, mu1, lv1 = self.nvp1(g2, mode=mode)
        else:
            raise ValueError(mode)
        return [g1, g2], [mu1, mu2], [lv1, lv2]
