"""Shared helpers for the test-suite (golden loading, tolerances)."""
import ast
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
GOLDEN_CASES = ['small_free_learned', 'small_freevar_global', 'small_fixed_learned', 'small_wide_learned',
                'small_mid_global', 'small_c3_freevar']


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
        self.meta = ast.literal_eval(str(self.z['meta']))

    def sd(self, dtype=torch.float64, device='cpu'):
        out = {}
        for k in self.z.files:
            if k.startswith('sd/'):
                t = torch.from_numpy(self.z[k])
                if t.is_floating_point():
                    t = t.to(dtype)
                out[k[3:]] = t.to(device)
        return out

    def t(self, key, dtype=torch.float64, device='cpu'):
        t = torch.from_numpy(np.asarray(self.z[key]))
        if t.is_floating_point():
            t = t.to(dtype)
        return t.to(device)

    def keys(self, prefix):
        return [k[len(prefix):] for k in self.z.files if k.startswith(prefix)]


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def max_rel(a, b, floor=1e-12):
    a = a.double()
    b = b.double()
    return float(((a - b).abs() / b.abs().clamp_min(floor)).max())


BASE_CFG = dict(
    train_mode='p_rnvp_mc_g_rnvp_vae', util_mode='training', deterministic=False,
    pc_enc_init_n_channels=3, pc_enc_init_n_features=8, pc_enc_n_features=[8, 16],
    g_prior_n_flows=1, g_prior_n_features=8, g_posterior_n_layers=1,
    p_latent_space_size=3, p_prior_n_layers=1, p_decoder_base_var=-3.9551,
    pnll_weight=1.0, gnll_weight=1.0, gent_weight=1.0,
)


def golden_cfg(gd):
    cfg = dict(BASE_CFG)
    cfg.update(gd.meta)
    return cfg


def build_dropin(gd, device='cpu'):
    """Our drop-in Flow_Mixture_Model carrying the golden (reference-generated) state_dict."""
    from go_with_the_flows_b200.networks.flow_mixture import Flow_Mixture_Model
    model = Flow_Mixture_Model(**golden_cfg(gd))
    model.load_state_dict(gd.sd(torch.float32), strict=True)
    return model.to(device)


def nll_err(a, b):
    """Per-point log-likelihood error with a mixed tolerance: |a-b| / max(|b|, 1).  A pure relative
    error is ill-defined where a point's NLL crosses zero (the 'fixed'-base golden case)."""
    a = a.double()
    b = b.double()
    return float(((a - b).abs() / b.abs().clamp_min(1.0)).max())
