"""Decoders.  Mirrors lib/networks/decoders.py: GlobalRNVPDecoder :7-38 (latent prior flow,
PyTorch) and LocalCondRNVPDecoder :41-79 (the per-point coupling stack, CUDA)."""
import torch.nn as nn

from .flows import CondRealNVPFlow3DTriple, RealNVPFlowCouple
from ..flowstack import FlowStack, run_module_stack


class GlobalRNVPDecoder(nn.Module):
    def __init__(self, n_flows, n_features, g_n_features, weight_std=0.01):
        super().__init__()
        self.n_flows = n_flows
        self.n_features = n_features
        self.g_n_features = g_n_features
        self.weight_std = weight_std
        self.flows = nn.ModuleList([
            RealNVPFlowCouple(n_features, g_n_features, weight_std=weight_std, pattern=i % 2)
            for i in range(n_flows)])

    def forward(self, g, mode='direct'):
        gs, mus, logvars = [], [], []
        order = range(self.n_flows) if mode == 'direct' else range(self.n_flows - 1, -1, -1)
        cur = g
        for i in order:
            a, b, c = self.flows[i](cur, mode=mode)
            if mode == 'direct':
                gs, mus, logvars = gs + a, mus + b, logvars + c
                cur = gs[-1]
            else:
                gs, mus, logvars = a + gs, b + mus, c + logvars
                cur = gs[0]
        return gs, mus, logvars


class LocalCondRNVPDecoder(nn.Module):
    """n_flows triples of conditional coupling layers, pattern i % 2 (decoders.py:49-52)."""

    def __init__(self, n_flows, f_n_features, g_n_features, weight_std=0.01):
        super().__init__()
        self.n_flows = n_flows
        self.f_n_features = f_n_features
        self.g_n_features = g_n_features
        self.weight_std = weight_std
        self.flows = nn.ModuleList([
            CondRealNVPFlow3DTriple(f_n_features, g_n_features, weight_std=weight_std, pattern=i % 2)
            for i in range(n_flows)])
        self._stack = None

    @staticmethod
    def get_param_count(n_flows, f_n_features, g_n_features):
        per_layer = 18 * f_n_features + 4 * f_n_features * g_n_features + 6 * f_n_features ** 2
        return n_flows * 3 * per_layer

    def coupling_layers(self):
        """The 3*n_flows coupling layers in DIRECT order."""
        return [m for t in self.flows for m in t.coupling_layers()]

    def forward(self, p, g, mode='direct'):
        """-> (ps, mus, logvars), 3*n_flows tensors each, indexed by layer in direct order:
        direct -> ps[-1] is the data-space sample; inverse -> ps[0] is the base-space sample
        (decoders.py:65-77)."""
        if self._stack is None:
            self._stack = FlowStack([self.coupling_layers()], own_storage=False)
        return run_module_stack(self._stack, p, g, mode, self.training)
