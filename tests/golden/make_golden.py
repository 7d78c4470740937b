"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports /root/reference/lib/networks (read-only), builds small mixture-of-flows models
with the reference constructors, perturbs BN statistics / last-layer weights away from
their near-identity init so every term of the math is exercised, runs the reference's own
`decode` + `FlowMixtureNLL` (train-mode BN and eval-mode BN, fp64 for a tight pin) and its
eval-mode sampling branch, and stores state_dict + inputs + outputs as .npz fixtures.
The GPU box has no /root/reference: tests only read the committed .npz files.
"""
import os
import sys

import numpy as np
import torch

REF = '/root/reference'
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

from lib.networks.flow_mixture import Flow_Mixture_Model  # noqa: E402
from lib.networks.losses import FlowMixtureNLL             # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

BASE_CFG = dict(
    train_mode='p_rnvp_mc_g_rnvp_vae', util_mode='training', deterministic=False,
    pc_enc_init_n_channels=3, pc_enc_init_n_features=8, pc_enc_n_features=[8, 16],
    g_prior_n_flows=1, g_prior_n_features=8, g_posterior_n_layers=1,
    p_latent_space_size=3, p_prior_n_layers=1, p_decoder_base_var=-3.9551,
    pnll_weight=1.0, gnll_weight=1.0, gent_weight=1.0,
)

CASES = {
    # name: (overrides, B, N)
    'small_free_learned': (dict(n_components=3, params_reduce_mode='none', weights_type='learned_weights',
                                p_decoder_n_flows=2, p_decoder_n_features=8, g_latent_space_size=16,
                                p_decoder_base_type='free'), 3, 50),
    'small_freevar_global': (dict(n_components=2, params_reduce_mode='depth_and_feature',
                                  weights_type='global_weights',
                                  p_decoder_n_flows=4, p_decoder_n_features=16, g_latent_space_size=32,
                                  p_decoder_base_type='freevar'), 4, 33),
    'small_fixed_learned': (dict(n_components=4, params_reduce_mode='none', weights_type='learned_weights',
                                 p_decoder_n_flows=1, p_decoder_n_features=5, g_latent_space_size=8,
                                 p_decoder_base_type='fixed'), 2, 64),
    # feature widths beyond the tcgen05 tile budget (F > 39: the mma.sync engine, padded to 48) and an odd
    # width padded to 32
    'small_wide_learned': (dict(n_components=2, params_reduce_mode='none', weights_type='learned_weights',
                                p_decoder_n_flows=1, p_decoder_n_features=44, g_latent_space_size=16,
                                p_decoder_base_type='free'), 2, 40),
    'small_mid_global': (dict(n_components=3, params_reduce_mode='none', weights_type='global_weights',
                              p_decoder_n_flows=1, p_decoder_n_features=27, g_latent_space_size=8,
                              p_decoder_base_type='freevar'), 3, 37),
    # the feature width of config_autoencoding.yaml / config_SVR.yaml (F = 33: FP = 36 on the FMA engine, 40 on the
    # tensor-core engines), freevar base, learned weights, ragged N spanning two 128-point tiles
    'small_c3_freevar': (dict(n_components=3, params_reduce_mode='none', weights_type='learned_weights',
                              p_decoder_n_flows=1, p_decoder_n_features=33, g_latent_space_size=16,
                              p_decoder_base_type='freevar', p_decoder_base_var=-3.596), 3, 150),
}


def build(overrides, seed):
    cfg = dict(BASE_CFG)
    cfg.update(overrides)
    torch.manual_seed(seed)
    model = Flow_Mixture_Model(**cfg).double()
    gen = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, t in model.state_dict().items():
            if name.endswith('running_mean'):
                t.copy_(0.3 * torch.randn(t.shape, generator=gen, dtype=t.dtype))
            elif name.endswith('running_var'):
                t.copy_(0.5 + torch.rand(t.shape, generator=gen, dtype=t.dtype))
            elif name.endswith('_bn.weight'):
                t.copy_(1.0 + 0.3 * torch.randn(t.shape, generator=gen, dtype=t.dtype))
            elif name.endswith('_bn.bias'):
                t.copy_(0.2 * torch.randn(t.shape, generator=gen, dtype=t.dtype))
            elif ('sd2.' in name or '_film_w1.' in name or '_film_b1.' in name) and 'pc_decoder' in name:
                # last layers start at N(0, 0.01^2) / 0 => near-identity flow; wake them up
                t.copy_(0.25 * torch.randn(t.shape, generator=gen, dtype=t.dtype))
            elif name == 'mixture_weights_logits':
                t.copy_(torch.randn(t.shape, generator=gen, dtype=t.dtype))
            elif name.startswith('mixture_weights_encoder.mus') or name.startswith('p_prior.mus') \
                    or name.startswith('p_prior.logvars'):
                t.copy_(0.3 * torch.randn(t.shape, generator=gen, dtype=t.dtype))
    return cfg, model


def point_nll_from_lists(output_decoder, logits):
    """Per-point quantities recomputed from the reference's own list outputs (losses.py:112-128)."""
    logw = logits - torch.logsumexp(logits, -1, keepdim=True)
    logps, zs = [], []
    for od in output_decoder:
        z = od['p_prior_samples'][0]
        S = sum(od['p_prior_logvars'])
        quad = (z - od['p_prior_mus'][0]) ** 2 / torch.exp(od['p_prior_logvars'][0])
        logps.append(-0.5 * ((S + quad).sum(1) + 3 * np.log(2 * np.pi)))
        zs.append(z)
    logp = torch.stack(logps, 2)
    nll = -torch.logsumexp(logp + logw.unsqueeze(1), -1)
    return nll, logp, torch.stack(zs, 1)


def run_nll(model, p, g, training, warmup=False):
    model.mode = 'training'
    model.train(training)
    sd_before = {k: v.clone() for k, v in model.state_dict().items()}
    p = p.clone().requires_grad_(True)
    g = g.clone().requires_grad_(True)
    model.zero_grad()
    out_dec, logits = model.decode(p, g, p.shape[2], False, warmup)
    pnll = FlowMixtureNLL()(out_dec, logits)
    pnll.backward()
    nll, logp, z = point_nll_from_lists(out_dec, logits)
    res = dict(pnll=pnll.detach(), nll=nll.detach(), logp=logp.detach(), z=z.detach(),
               logits=logits.detach(), dp=p.grad.clone(), dg=g.grad.clone())
    grads = {k: (v.grad.clone() if v.grad is not None else torch.zeros_like(v))
             for k, v in model.named_parameters()
             if k.startswith(('pc_decoder', 'p_prior', 'mixture_weights'))}
    sd_after = {k: v.clone() for k, v in model.state_dict().items()}
    # restore buffers so the next run starts from the same state
    model.load_state_dict(sd_before)
    changed = {k: v for k, v in sd_after.items()
               if ('running_' in k or 'num_batches' in k) and not torch.equal(v, sd_before[k])}
    return res, grads, changed


def run_sample(model, g, idx, eps):
    """Drive the reference's eval branch (flow_mixture.py:141-177) shape by shape with OUR
    component indices and noise, by standing in for np.random.choice / randn_like."""
    import lib.networks.flow_mixture as fm
    model.mode = 'generating'
    model.eval()
    B, _, N = eps.shape
    outs, labels = [], []
    real_choice = fm.np.random.choice
    real_reparam = model.reparameterize
    with torch.no_grad():
        for b in range(B):
            state = {'t': 0}

            def fake_choice(a, size=None, p=None, _b=b):
                return idx[_b].copy()

            def fake_reparam(mu, logvar, _b=b):
                t = state['t']
                state['t'] += 1
                e = eps[_b:_b + 1][:, :, torch.as_tensor(idx[_b] == t)]
                return e * torch.exp(0.5 * logvar) + mu

            fm.np.random.choice = fake_choice
            model.reparameterize = fake_reparam
            try:
                s, lab, _ = model.decode(torch.zeros(1, 3, N, dtype=eps.dtype), g[b:b + 1], N, True, False)
            finally:
                fm.np.random.choice = real_choice
                model.reparameterize = real_reparam
            outs.append(s)
            labels.append(lab)
    model.mode = 'training'
    return torch.cat(outs), torch.cat(labels)


def to_np(d, prefix):
    return {prefix + k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def main():
    only = set(sys.argv[1:])          # optional: regenerate just the named cases (seeds depend on the case index)
    for ci, (name, (ov, B, N)) in enumerate(CASES.items()):
        if only and name not in only:
            continue
        cfg, model = build(ov, 100 + ci)
        gen = torch.Generator().manual_seed(7 + ci)
        G = cfg['g_latent_space_size']
        p = 0.4 * torch.randn(B, 3, N, generator=gen, dtype=torch.float64)
        g = 0.7 * torch.randn(B, G, generator=gen, dtype=torch.float64)
        blob = {}
        blob.update(to_np(model.state_dict(), 'sd/'))
        blob['in/p'] = p.numpy()
        blob['in/g'] = g.numpy()
        for tag, training in (('train', True), ('eval', False)):
            res, grads, changed = run_nll(model, p, g, training)
            blob.update(to_np(res, f'{tag}/'))
            blob.update(to_np(grads, f'{tag}/grad/'))
            blob.update(to_np(changed, f'{tag}/bn/'))
        # warm-up branch (global logits even for learned weights)
        res, _, _ = run_nll(model, p, g, False, warmup=True)
        blob['eval_warmup/pnll'] = res['pnll'].numpy()
        # sampling
        K = cfg['n_components']
        rs = np.random.RandomState(11 + ci)
        idx = rs.randint(0, K, size=(B, N)).astype(np.int32)
        idx[0, :] = 0 if K > 1 else 0          # one shape entirely on component 0 (others get 0 points)
        eps = torch.randn(B, 3, N, generator=gen, dtype=torch.float64)
        x, labels = run_sample(model, g, idx, eps)
        blob['sample/idx'] = idx
        blob['sample/eps'] = eps.numpy()
        blob['sample/x'] = x.numpy()
        blob['sample/labels'] = labels.numpy()
        meta = {k: v for k, v in cfg.items() if k in (
            'n_components', 'params_reduce_mode', 'weights_type', 'p_decoder_n_flows',
            'p_decoder_n_features', 'g_latent_space_size', 'p_decoder_base_type', 'p_decoder_base_var')}
        blob['meta'] = np.array(repr(meta))
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **blob)
        print(name, 'pnll train/eval', float(blob['train/pnll']), float(blob['eval/pnll']),
              'file KB', os.path.getsize(path) // 1024)


if __name__ == '__main__':
    main()
