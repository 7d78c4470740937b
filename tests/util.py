"""Shared helpers for the test-suite (golden loading, tolerances)."""
import ast
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
GOLDEN_CASES = ['small_free_learned', 'small_freevar_global', 'small_fixed_learned']


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
