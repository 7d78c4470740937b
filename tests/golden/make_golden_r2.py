"""Round-2 golden vectors from the UNMODIFIED reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_r2.py

Covers the entry points the CLIs actually call, on top of make_golden.py's decode-level cases:

  model_small.npz    `model(g_input, p_input)` (models.py:224-258) + `Flow_Mixture_Loss` (losses.py:159-173) +
                     backward, train mode, posterior noise injected: loss, pnll, gnll, gent, every gradient.
  oneflow_small.npz  `one_flow_decode` (models.py:153-207) of ONE decoder in training mode + `PointFlowNLL`
                     (losses.py:7-20) + backward: the per-module list API under autograd.
  svr_full.npz       `Flow_Mixture_SVR_Model` at config_SVR.yaml size (K=4, L=33, F=33, G=512, freevar, 2500
                     points), weights from `torch.manual_seed(0)` + the reference constructor (NOT stored: the
                     drop-in's seeded construction is bit-identical, tests/test_dropin_cpu.py), eval mode
                     'reconstruction': latent from synthetic images, samples for given component indices / noise.
  adam_steps.npz     the custom `Adam` (optimizers.py:15-76; AMSGrad, un-scaled weight decay) + `LRUpdater`,
                     five steps on three small tensors with stored gradients.
"""
import os
import sys

import numpy as np
import torch

REF = '/root/reference'
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from lib.networks.flow_mixture import Flow_Mixture_Model, Flow_Mixture_SVR_Model  # noqa: E402
from lib.networks.losses import Flow_Mixture_Loss, PointFlowNLL                   # noqa: E402
from lib.networks.optimizers import Adam, LRUpdater                               # noqa: E402

from tests.golden.make_golden import CASES, build, to_np                          # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def save(name, blob):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **blob)
    print(name, 'file KB', os.path.getsize(path) // 1024)


def model_small():
    """Whole-model forward + loss + backward on the 'small_free_learned' model (same seed => same state_dict as
    the decode-level golden of that name, re-stored here so the file stands alone)."""
    ov, B, N = CASES['small_free_learned']
    cfg, model = build(ov, 100)
    gen = torch.Generator().manual_seed(4242)
    G = cfg['g_latent_space_size']
    g_in = 0.4 * torch.randn(B, 3, N, generator=gen, dtype=torch.float64)       # batch['cloud']
    p_in = 0.4 * torch.randn(B, 3, N, generator=gen, dtype=torch.float64)       # batch['eval_cloud']
    eps = torch.randn(B, G, generator=gen, dtype=torch.float64)                 # posterior noise (models.py:107)
    model.mode = 'training'
    model.train()
    model.reparameterize = lambda mu, logvar: eps * torch.exp(0.5 * logvar) + mu
    blob = to_np(model.state_dict(), 'sd/')
    blob.update({'in/g_input': g_in.numpy(), 'in/p_input': p_in.numpy(), 'in/eps': eps.numpy()})
    loss_fn = Flow_Mixture_Loss(**cfg)
    for tag, warmup in (('train', False), ('warmup', True)):
        sd0 = {k: v.clone() for k, v in model.state_dict().items()}
        model.zero_grad()
        out_enc, out_dec, logits = model(g_in, p_in, images=None, n_sampled_points=None, labeled_samples=False,
                                         warmup=warmup)
        loss, pnll, gnll, gent = loss_fn(out_enc, out_dec, logits)
        loss.backward()
        blob.update(to_np(dict(loss=loss.detach(), pnll=pnll.detach(), gnll=gnll.detach(), gent=gent.detach(),
                               logits=logits.detach(), g_sample=out_enc['g_posterior_samples'].detach(),
                               g_prior_z=out_enc['g_prior_samples'][0].detach()), f'{tag}/'))
        grads = {k: (v.grad.clone() if v.grad is not None else torch.zeros_like(v)) for k, v in model.named_parameters()}
        blob.update(to_np(grads, f'{tag}/grad/'))
        sd1 = model.state_dict()
        changed = {k: v.clone() for k, v in sd1.items()
                   if ('running_' in k or 'num_batches' in k) and not torch.equal(v, sd0[k])}
        blob.update(to_np(changed, f'{tag}/bn/'))
        model.load_state_dict(sd0)
    blob['meta'] = np.array(repr({k: v for k, v in cfg.items() if not isinstance(v, (list, tuple)) or True}))
    save('model_small', blob)


def oneflow_small():
    ov, B, N = CASES['small_free_learned']
    cfg, model = build(ov, 100)
    gen = torch.Generator().manual_seed(777)
    G = cfg['g_latent_space_size']
    p = (0.4 * torch.randn(B, 3, N, generator=gen, dtype=torch.float64)).requires_grad_(True)
    g = (0.7 * torch.randn(B, G, generator=gen, dtype=torch.float64)).requires_grad_(True)
    model.mode = 'training'
    blob = {'in/p': p.detach().numpy(), 'in/g': g.detach().numpy()}
    for tag, training in (('train', True), ('eval', False)):
        sd0 = {k: v.clone() for k, v in model.state_dict().items()}
        model.train(training)
        model.zero_grad()
        p.grad = g.grad = None
        out = model.one_flow_decode(p, g, model.pc_decoder[1], N)
        nll = PointFlowNLL()(out)                       # (B,1,N)
        nll.sum(dim=2).mean().backward()
        blob.update(to_np(dict(nll=nll.detach(), z=out['p_prior_samples'][0].detach(),
                               logdet=sum(out['p_prior_logvars']).detach(), dp=p.grad.clone(), dg=g.grad.clone(),
                               n_entries=np.array(len(out['p_prior_samples']))), f'{tag}/'))
        grads = {k: v.grad.clone() for k, v in model.named_parameters() if v.grad is not None}
        blob.update(to_np(grads, f'{tag}/grad/'))
        model.load_state_dict(sd0)
    save('oneflow_small', blob)


def svr_full():
    import yaml
    cfg = yaml.safe_load(open(os.path.join(REF, 'configs', 'config_SVR.yaml')))
    cfg['weights_type'] = 'learned_weights'
    cfg['util_mode'] = 'training'
    torch.manual_seed(0)
    model = Flow_Mixture_SVR_Model(**cfg)
    model.mode = 'reconstruction'
    model.eval()
    B, N = 2, cfg['cloud_size']
    gen = torch.Generator().manual_seed(99)
    images = torch.randn(B, 4, 224, 224, generator=gen)
    with torch.no_grad():
        enc = model.encode(None, images)
        g = enc['g_prior_samples'][-1]
        logits = model.get_weights(g)
    K = cfg['n_components']
    rs = np.random.RandomState(5)
    idx = rs.randint(0, K, size=(B, N)).astype(np.int32)
    eps = torch.randn(B, 3, N, generator=gen, dtype=torch.float64)
    # fp64 sampling with the reference's own eval branch, driven shape by shape with our draws
    from tests.golden.make_golden import run_sample
    model64 = model.double()
    model64.mode = 'reconstruction'
    x, labels = run_sample(model64, g.double(), idx, eps)
    blob = {'in/image_seed': np.array(99), 'svr/g': g.numpy(), 'svr/img_features_mu0': enc['g_prior_mus'][0].numpy(),
            'svr/logits': logits.numpy(), 'sample/idx': idx, 'sample/eps': eps.numpy().astype(np.float32),
            'sample/x': x.numpy(), 'sample/labels': labels.numpy()}
    save('svr_full', blob)


def adam_steps():
    gen = torch.Generator().manual_seed(31)
    shapes = [(7, 5), (13,), (2, 3, 4)]
    params = [torch.nn.Parameter(torch.randn(s, generator=gen)) for s in shapes]
    blob = {f'p0/{i}': p.detach().numpy().copy() for i, p in enumerate(params)}
    opt = Adam(params, lr=2.56e-4, betas=(0.9, 0.995), weight_decay=1e-4, amsgrad=True)
    sched = LRUpdater(4, cycle_length=2, min_lr=1e-5, max_lr=3e-3, beta1=0.9, min_beta2=0.99, max_beta2=0.999)
    for step in range(5):
        sched(opt, step // 4, step % 4)
        for i, p in enumerate(params):
            p.grad = torch.randn(p.shape, generator=gen) * (10.0 ** (step - 2))
            blob[f'g{step}/{i}'] = p.grad.numpy().copy()
        blob[f'lr{step}'] = np.array([opt.param_groups[0]['lr'], opt.param_groups[0]['betas'][1]])
        opt.step()
        for i, p in enumerate(params):
            blob[f'p{step + 1}/{i}'] = p.detach().numpy().copy()
    save('adam_steps', blob)


if __name__ == '__main__':
    which = set(sys.argv[1:]) or {'model_small', 'oneflow_small', 'svr_full', 'adam_steps'}
    for name in sorted(which):
        globals()[name]()
