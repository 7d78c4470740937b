"""CPU oracle for the mixture-of-flows hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional (module-free) restatement, in plain torch CPU ops, of the
reference's per-point mixture-of-conditional-RealNVP log-likelihood and its
sampling pass.  It is the checker the CUDA kernels are compared against and the
CPU baseline `bench.py --impl reference` times.  Nothing in the product package
(`go_with_the_flows_b200/`) may import it; only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s CPU-baseline legs do.

Parity pin: the reference ships no tests or golden vectors for this path
(SURVEY.md §4), so the oracle is pinned against the reference's own modules run
in the build container: `tests/golden/make_golden.py` imports
`/root/reference/lib/networks`, runs it under fixed seeds and commits the
input/output vectors under `tests/golden/`; `tests/test_oracle_golden.py`
replays them through this file.

Parameters are addressed by the reference's state_dict keys,
e.g. ``pc_decoder.0.flows.3.nvp2.T_mu_0.mu_sd1.weight`` so any reference
checkpoint can be fed in unchanged.  dtype follows the tensors handed in, so a
``.double()`` state_dict gives the fp64 oracle used as the gradient yardstick.

Reference lines restated (all under /root/reference/lib/networks/):
  layers.py:40-45     SharedDot.forward          -> _shared_dot
  flows.py:95-117     CondRealNVPFlow3D.forward  -> coupling_layer
  flows.py:129-160    Triple order / patterns    -> layer_plan
  decoders.py:61-79   LocalCondRNVPDecoder       -> decoder_stack
  encoders.py:72-89   FeatureEncoder / Weights   -> feature_encoder
  models.py:153-207   one_flow_decode            -> base_gaussian, component_logp, sample
  flow_mixture.py:104-179  get_weights / decode  -> mixture_logits, sample
  losses.py:88-137    FlowMixtureNLL             -> mixture_nll
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
LOG_2PI = math.log(2.0 * math.pi)

# ----------------------------------------------------------------------------
# structure helpers
# ----------------------------------------------------------------------------

# flows.py:129-148 -- warp sets of nvp1..nvp3 for the two triple patterns
_WARP_SETS = {
    0: ([0], [1], [2]),
    1: ([0, 1], [0, 2], [1, 2]),
}


def layer_plan(n_flows: int) -> List[Tuple[int, int, List[int], List[int]]]:
    """(triple index, nvp number 1..3, warp dims, keep dims) in DIRECT order.

    decoders.py:49-52 gives triple i pattern i % 2; flows.py:151-154 runs
    nvp1 -> nvp2 -> nvp3 inside a triple.  The inverse pass walks this list
    backwards (decoders.py:72-77, flows.py:155-158).
    """
    plan = []
    for i in range(n_flows):
        for n, warp in enumerate(_WARP_SETS[i % 2]):
            keep = [d for d in (0, 1, 2) if d not in warp]
            plan.append((i, n + 1, list(warp), keep))
    return plan


@dataclass
class DecoderDims:
    n_components: int
    n_flows: int      # triples per component
    n_features: int   # F
    g_features: int   # G

    @property
    def n_layers(self) -> int:
        return 3 * self.n_flows


def infer_dims(sd: Dict[str, torch.Tensor], prefix: str = 'pc_decoder.') -> DecoderDims:
    """Recover K, n_flows, F, G from state_dict keys/shapes."""
    comps, flows = set(), set()
    pat = re.compile(re.escape(prefix) + r'(\d+)\.flows\.(\d+)\.nvp1\.T_mu_0\.mu_sd1\.weight$')
    for k in sd:
        m = pat.match(k)
        if m:
            comps.add(int(m.group(1)))
            flows.add(int(m.group(2)))
    if not comps:
        raise KeyError('no mixture decoder found under prefix %r' % prefix)
    w = sd[prefix + '0.flows.0.nvp1.T_mu_0_cond_w.mu_sd1_film_w0.weight']
    return DecoderDims(len(comps), len(flows), int(w.shape[0]), int(w.shape[1]))


class BNUpdates(dict):
    """Collects the running-stat updates a train-mode pass would apply.

    key -> new tensor, for `.running_mean`, `.running_var`; `.num_batches_tracked`
    keys map to the increment count (p_prior is called K times per forward in the
    reference, models.py:171 via flow_mixture.py:163-166, so its BN advances K
    times per step)."""


# ----------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------

def _shared_dot(weight: torch.Tensor, x: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    # layers.py:41-44: weight (1,out,in) @ x (B,1,in,N) -> (B,1,out,N)
    out = torch.matmul(weight, x.unsqueeze(1)).squeeze(1)
    if bias is not None:
        out = out + bias.unsqueeze(0).unsqueeze(3).squeeze(0)
    return out


def _batch_norm(x, sd, key, training, affine=True, upd: Optional[BNUpdates] = None):
    """nn.BatchNorm1d semantics (biased var to normalise, unbiased into running_var)."""
    rm, rv = sd[key + '.running_mean'], sd[key + '.running_var']
    w = sd[key + '.weight'] if affine else None
    b = sd[key + '.bias'] if affine else None
    if training:
        dims = (0,) if x.dim() == 2 else (0, 2)
        n = x.numel() // x.shape[1]
        mean = x.mean(dims)
        var = x.var(dims, unbiased=False)
        if upd is not None:
            with torch.no_grad():
                # repeated application (p_prior): start from the latest value
                cur_m = upd.get(key + '.running_mean', rm)
                cur_v = upd.get(key + '.running_var', rv)
                upd[key + '.running_mean'] = (1 - BN_MOMENTUM) * cur_m + BN_MOMENTUM * mean
                upd[key + '.running_var'] = (1 - BN_MOMENTUM) * cur_v + BN_MOMENTUM * var * (n / max(n - 1, 1))
                upd[key + '.num_batches_tracked'] = upd.get(key + '.num_batches_tracked', 0) + 1
    else:
        mean, var = rm, rv
    shape = (1, -1) if x.dim() == 2 else (1, -1, 1)
    y = (x - mean.view(shape)) / torch.sqrt(var.view(shape) + BN_EPS)
    if affine:
        y = y * w.view(shape) + b.view(shape)
    return y


def _swish(x):
    return x * torch.sigmoid(x)


def _cond_net(g, sd, base, stem, training, upd):
    # flows.py:33-45 -- Linear(G->F, no bias) -> BN -> Swish -> Linear(F->F)+bias
    h = F.linear(g, sd[f'{base}.{stem}0.weight'])
    h = _batch_norm(h, sd, f'{base}.{stem}0_bn', training, True, upd)
    h = _swish(h)
    return F.linear(h, sd[f'{base}.{stem}1.weight'], sd[f'{base}.{stem}1.bias'])


def _point_net(x_keep, g, sd, lp, X, training, upd):
    """One of the two per-point MLPs (X = 'mu' | 'logvar'); flows.py:25-50,99-107."""
    eps = sd[lp + 'eps']
    h = _shared_dot(sd[f'{lp}T_{X}_0.{X}_sd0.weight'], x_keep)
    h = _batch_norm(h, sd, f'{lp}T_{X}_0.{X}_sd0_bn', training, True, upd)
    h = torch.relu(h)
    h = _shared_dot(sd[f'{lp}T_{X}_0.{X}_sd1.weight'], h)
    h = _batch_norm(h, sd, f'{lp}T_{X}_0.{X}_sd1_bn', training, False, upd)
    s = eps + torch.exp(_cond_net(g, sd, f'{lp}T_{X}_0_cond_w', f'{X}_sd1_film_w', training, upd))
    t = _cond_net(g, sd, f'{lp}T_{X}_0_cond_b', f'{X}_sd1_film_b', training, upd)
    h = torch.relu(s.unsqueeze(2) * h + t.unsqueeze(2))
    return _shared_dot(sd[f'{lp}T_{X}_1.{X}_sd2.weight'], h, sd[f'{lp}T_{X}_1.{X}_sd2.bias'])


def coupling_layer(p, g, sd, lp, warp, keep, mode, training, upd=None):
    """flows.py:95-117.  p (B,3,N), g (B,G) -> (p_out, mu, logvar) each (B,3,N)."""
    eps = sd[lp + 'eps']
    x_keep = p[:, keep, :].contiguous()
    # the reference evaluates the logvar net first (flows.py:99) then mu (:104)
    o_lv = _point_net(x_keep, g, sd, lp, 'logvar', training, upd)
    o_mu = _point_net(x_keep, g, sd, lp, 'mu', training, upd)
    logvar = torch.zeros_like(p)
    mu = torch.zeros_like(p)
    logvar[:, warp, :] = o_lv / (1.0 + o_lv.abs())          # softsign
    mu[:, warp, :] = o_mu
    scale = torch.sqrt(eps + torch.exp(logvar))               # on ALL dims: keep dims get sqrt(eps+1)
    if mode == 'direct':
        p_out = scale * p + mu
    elif mode == 'inverse':
        p_out = (p - mu) / scale
    else:
        raise ValueError(mode)
    return p_out, mu, logvar


def decoder_stack(p, g, sd, prefix, n_flows, mode, training, upd=None):
    """decoders.py:61-79.  Returns (ps, mus, logvars) lists of 3*n_flows tensors with
    the reference ordering: direct -> index -1 is the data-space sample; inverse ->
    index 0 is the base-space sample."""
    plan = layer_plan(n_flows)
    order = plan if mode == 'direct' else plan[::-1]
    ps, mus, lvs = [], [], []
    cur = p
    for (i, n, warp, keep) in order:
        lp = f'{prefix}flows.{i}.nvp{n}.'
        cur, mu, lv = coupling_layer(cur, g, sd, lp, warp, keep, mode, training, upd)
        if mode == 'direct':
            ps.append(cur); mus.append(mu); lvs.append(lv)
        else:
            ps.insert(0, cur); mus.insert(0, mu); lvs.insert(0, lv)
    return ps, mus, lvs


def feature_encoder(x, sd, prefix, n_layers, deterministic, training, upd=None):
    """encoders.py:31-83."""
    h = x
    for i in range(n_layers):
        h = F.linear(h, sd[f'{prefix}features.mlp{i}.weight'])
        h = _batch_norm(h, sd, f'{prefix}features.mlp{i}_bn', training, True, upd)
        h = _swish(h)
    mu = F.linear(h, sd[f'{prefix}mus.mu_mlp0.weight'], sd[f'{prefix}mus.mu_mlp0.bias'])
    if deterministic:
        return mu
    lv = F.linear(h, sd[f'{prefix}logvars.logvar_mlp0.weight'], sd[f'{prefix}logvars.logvar_mlp0.bias'])
    return mu, lv


def base_gaussian(g, sd, base_type, training, upd=None, base_var=None, n_layers=1):
    """models.py:169-193 -> (mu_base, lv_base) each (B,3)."""
    B = g.shape[0]
    if base_type == 'free':
        return feature_encoder(g, sd, 'p_prior.', n_layers, False, training, upd)
    if base_type == 'freevar':
        lv = feature_encoder(g, sd, 'p_prior.', n_layers, True, training, upd)
        return torch.zeros_like(lv), lv
    if base_type == 'fixed':
        # models.py:90-92 -- constants live in (1,3,1) buffers (fp32-rounded base_var)
        if 'p_prior_logvar' in sd:
            return (sd['p_prior_mus'].view(1, 3).expand(B, 3),
                    sd['p_prior_logvar'].view(1, 3).expand(B, 3))
        mu = g.new_zeros(B, 3)
        return mu, mu + float(base_var)
    raise ValueError(base_type)


def mixture_logits(g, sd, weights_type, warmup, training, upd=None):
    """flow_mixture.py:104-120 -> (B,K) un-normalised log-weights."""
    if warmup or weights_type == 'global_weights':
        w = sd['mixture_weights_logits']
        return w.unsqueeze(0).expand(g.shape[0], w.shape[0])
    mus = feature_encoder(g, sd, 'mixture_weights_encoder.', 3, True, training, upd)
    return F.log_softmax(mus, dim=1)


# ----------------------------------------------------------------------------
# the NLL pass  (reference mode='inverse')
# ----------------------------------------------------------------------------

def component_logp(p, g, sd, j, n_flows, mu_b, lv_b, training, upd=None):
    """log N(z_j; base) - sum logvar  for component j; losses.py:112-122.  -> (B,N), z, S"""
    ps, _, lvs = decoder_stack(p, g, sd, f'pc_decoder.{j}.', n_flows, 'inverse', training, upd)
    z = ps[0]
    S = lv_b.unsqueeze(2)
    for lv in lvs:
        S = S + lv
    quad = (z - mu_b.unsqueeze(2)) ** 2 / torch.exp(lv_b.unsqueeze(2))
    logp = -0.5 * ((S + quad).sum(1) + 3.0 * LOG_2PI)
    return logp, z, S


def mixture_nll(p, g, sd, base_type='free', weights_type='learned_weights', warmup=False,
                training=True, upd=None, base_var=None, logits=None, base=None):
    """Full hot path: decode (flow_mixture.py:122-166, training branch) + FlowMixtureNLL.

    Returns dict(pnll scalar, nll (B,N), logp (B,N,K), logits (B,K), mu_base, lv_base).
    `logits` / `base` may be injected to test the flow stack in isolation.
    """
    dims = infer_dims(sd)
    K = dims.n_components
    if logits is None:
        logits = mixture_logits(g, sd, weights_type, warmup, training, upd)
    logps = []
    mu_b = lv_b = None
    for j in range(K):
        # models.py:171 -- p_prior is re-evaluated for every component
        if base is None:
            mu_b, lv_b = base_gaussian(g, sd, base_type, training, upd, base_var)
        else:
            mu_b, lv_b = base
        lp, _, _ = component_logp(p, g, sd, j, dims.n_flows, mu_b, lv_b, training, upd)
        logps.append(lp)
    logp = torch.stack(logps, dim=2)                              # (B,N,K)
    logw = logits - torch.logsumexp(logits, dim=-1, keepdim=True)  # losses.py:101-104
    nll = -torch.logsumexp(logp + logw.unsqueeze(1), dim=-1)      # (B,N)
    pnll = nll.sum(1).mean()                                      # losses.py:129-135
    return dict(pnll=pnll, nll=nll, logp=logp, logits=logits, mu_base=mu_b, lv_base=lv_b)


# ----------------------------------------------------------------------------
# sampling  (reference mode='direct')
# ----------------------------------------------------------------------------

PHILOX_M0 = 0xD2511F53
PHILOX_M1 = 0xCD9E8D57
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Philox4x32-10 (Salmon et al. 2011, Random123).  ctr (...,4) uint32, key (...,2) uint32."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0 = key[..., 0].astype(np.uint64)
    k1 = key[..., 1].astype(np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(PHILOX_M0) * c[0]
        p1 = np.uint64(PHILOX_M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(PHILOX_W0)) & mask
        k1 = (k1 + np.uint64(PHILOX_W1)) & mask
    return np.stack([x.astype(np.uint32) for x in c], axis=-1)


def u01(x: np.ndarray) -> np.ndarray:
    """uint32 -> float32 in [0,1): (x >> 8) * 2^-24, exact in fp32."""
    return ((x >> np.uint32(8)).astype(np.float32)) * np.float32(2.0 ** -24)


def sample_streams(seed: int, stream: int, B: int, N: int):
    """The per-point random numbers the sampling kernel draws.

    counter = (n, b, call, 0), key = (seed & 0xffffffff, stream ^ (seed >> 32)).  call 0 -> words
    (r0,r1,r2,r3), call 1 -> (r4,..).  u_comp = u01(r0);  Box-Muller on
    (r1,r2) -> eps0, eps1 and on (r3,r4) -> eps2 (cosine branch only).
    Returns u_comp (B,N) float32 and the raw uint32 words (B,N,8).
    """
    n = np.arange(N, dtype=np.uint32)[None, :].repeat(B, 0)
    b = np.arange(B, dtype=np.uint32)[:, None].repeat(N, 1)
    key = np.zeros((B, N, 2), dtype=np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((stream ^ (seed >> 32)) & 0xFFFFFFFF)
    words = []
    for call in (0, 1):
        ctr = np.stack([n, b, np.full_like(n, call), np.zeros_like(n)], axis=-1)
        words.append(philox4x32_10(ctr, key))
    w = np.concatenate(words, axis=-1)
    return u01(w[..., 0]), w


def box_muller(w: np.ndarray) -> np.ndarray:
    """(B,N,8) uint32 words -> (B,3,N) float32 standard normals (kernel convention).

    u = (r >> 8 + 0.5) * 2^-24 in (0,1) for the radius so log() never sees 0."""
    def uo(x):
        return ((x >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)
    two_pi = np.float32(2.0 * math.pi)
    r01 = np.sqrt(np.float32(-2.0) * np.log(uo(w[..., 1])))
    th01 = two_pi * u01(w[..., 2])
    r2 = np.sqrt(np.float32(-2.0) * np.log(uo(w[..., 3])))
    th2 = two_pi * u01(w[..., 4])
    e0 = r01 * np.cos(th01)
    e1 = r01 * np.sin(th01)
    e2 = r2 * np.cos(th2)
    return np.stack([e0, e1, e2], axis=1).astype(np.float32)


def mixture_cdf(logits_row: np.ndarray) -> np.ndarray:
    """flow_mixture.py:149-150 probs, then the float64 inclusive cumsum numpy's
    `choice` builds; stored as float32 with the last entry forced to 1.

    The fp32 exponential is taken as the correctly rounded one, fp32(exp(fp64(x))), and the fp32 sum runs left to
    right: numpy's own float32 `exp` may differ from it in the last bit depending on the SIMD path it takes, and
    a rule that both this restatement and the CUDA kernel (k_mixture_cdf) can reproduce bit for bit is what makes
    the component assignment bit-exact.  A one-ulp change of a probability moves a cdf boundary by ~6e-8."""
    e = np.exp(logits_row.astype(np.float32).astype(np.float64)).astype(np.float32)
    tot = np.float32(0.0)
    for v in e:
        tot = np.float32(tot + v)
    probs = e / tot
    cdf = np.cumsum(probs.astype(np.float64))
    cdf /= cdf[-1]
    cdf = cdf.astype(np.float32)
    cdf[-1] = np.float32(1.0)
    return cdf


def component_index(cdf: np.ndarray, u: np.ndarray) -> np.ndarray:
    """np.random.choice's rule (flow_mixture.py:153): cdf.searchsorted(u, side='right')."""
    idx = np.searchsorted(cdf, u, side='right')
    return np.minimum(idx, len(cdf) - 1).astype(np.int32)


def sample(g, sd, idx, eps, base_type='freevar', base_var=None, base=None):
    """Eval-mode sampling given component indices idx (B,N) int and noise eps (B,3,N).

    flow_mixture.py:141-177 + models.py:199-203, lifted from B==1 to a batch: point n of
    shape b is produced by component idx[b,n] from z = mu_base + exp(lv_base/2)*eps.
    Running every component on every point and selecting is arithmetically identical
    to the reference's gather/scatter because points are independent in eval mode.
    Returns x (B,3,N) and labels (B,N) = idx+1 (flow_mixture.py:176).
    """
    dims = infer_dims(sd)
    if base is None:
        mu_b, lv_b = base_gaussian(g, sd, base_type, False, None, base_var)
    else:
        mu_b, lv_b = base
    z = mu_b.unsqueeze(2) + torch.exp(0.5 * lv_b).unsqueeze(2) * eps
    idx_t = torch.as_tensor(idx, dtype=torch.long)
    out = torch.zeros_like(z)
    for j in range(dims.n_components):
        ps, _, _ = decoder_stack(z, g, sd, f'pc_decoder.{j}.', dims.n_flows, 'direct', False)
        out = torch.where((idx_t == j).unsqueeze(1), ps[-1], out)
    return out, (idx_t + 1).to(z.dtype), z
