"""CPU tests of the host-side modules around the hot path against golden vectors generated from the unmodified
reference (tests/golden/make_golden_r2.py): optimizer + schedule, encoders / prior flow / latent-side losses,
the SVR model's image branch -- and, when /root/reference is present (the build container), directly against
the reference's own modules."""
import ast
import os
import sys

import numpy as np
import pytest
import torch

from tests.util import GOLDEN_DIR, max_rel, rel_l2

REF = '/root/reference'
HAVE_REF = os.path.isdir(os.path.join(REF, 'lib', 'networks'))


def _npz(name):
    return np.load(os.path.join(GOLDEN_DIR, name + '.npz'))


def _ref_modules():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import lib.networks.encoders as enc
    import lib.networks.flow_mixture as fm
    import lib.networks.losses as losses
    import lib.networks.optimizers as opt
    return enc, fm, losses, opt


# ------------------------------------------------------------------------------------------ optimizer
def test_adam_and_lrupdater_reproduce_the_reference_steps():
    """optimizers.py:15-97: AMSGrad, bias corrections, weight decay added un-scaled by lr, cosine schedule on lr
    and beta2 -- five steps, parameter for parameter."""
    from go_with_the_flows_b200.networks.optimizers import Adam, LRUpdater
    z = _npz('adam_steps')
    params = [torch.nn.Parameter(torch.from_numpy(z[f'p0/{i}']).clone()) for i in range(3)]
    opt = Adam(params, lr=2.56e-4, betas=(0.9, 0.995), weight_decay=1e-4, amsgrad=True)
    sched = LRUpdater(4, cycle_length=2, min_lr=1e-5, max_lr=3e-3, beta1=0.9, min_beta2=0.99, max_beta2=0.999)
    for step in range(5):
        sched(opt, step // 4, step % 4)
        assert np.allclose([opt.param_groups[0]['lr'], opt.param_groups[0]['betas'][1]], z[f'lr{step}'], rtol=1e-12)
        for i, p in enumerate(params):
            p.grad = torch.from_numpy(z[f'g{step}/{i}']).clone()
        opt.step()
        for i, p in enumerate(params):
            want = torch.from_numpy(z[f'p{step + 1}/{i}'])
            assert float((p.detach() - want).abs().max()) <= 2e-7 * max(1.0, float(want.abs().max())), (step, i)
    st = opt.state_dict()['state']
    assert all(int(st[i]['step']) == 5 and 'max_exp_avg_sq' in st[i] for i in range(3))


@pytest.mark.skipif(not HAVE_REF, reason='needs /root/reference (build container only)')
def test_adam_without_amsgrad_or_decay_matches_reference_module():
    from go_with_the_flows_b200.networks.optimizers import Adam
    _, _, _, ropt = _ref_modules()
    gen = torch.Generator().manual_seed(3)
    a = [torch.nn.Parameter(torch.randn(11, 3, generator=gen)) for _ in range(2)]
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    oa, ob = Adam(a, lr=1e-3), ropt.Adam(b, lr=1e-3)
    for _ in range(4):
        for p, q in zip(a, b):
            p.grad = torch.randn(p.shape, generator=gen)
            q.grad = p.grad.clone()
        oa.step()
        ob.step()
    for p, q in zip(a, b):
        assert float((p - q).abs().max()) < 1e-7


# ------------------------------------------------------------------------------------------ latent side of model()
def _model_small(device='cpu'):
    from go_with_the_flows_b200.networks.flow_mixture import Flow_Mixture_Model
    z = _npz('model_small')
    cfg = ast.literal_eval(str(z['meta']))
    model = Flow_Mixture_Model(**cfg)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd/')}
    sd = {k: (v.float() if v.is_floating_point() else v) for k, v in sd.items()}
    model.load_state_dict(sd, strict=True)
    return z, cfg, model.to(device)


def test_encoder_prior_flow_and_latent_losses_match_reference():
    """models.py:111-151 `encode` (PointNet encoder -> max-pool -> posterior -> reparameterize -> prior flow,
    inverse) and the latent-side terms of Flow_Mixture_Loss (losses.py:23-39,168-169), train mode, injected
    posterior noise.  Pure PyTorch modules: run on the CPU."""
    from go_with_the_flows_b200.networks.losses import GaussianEntropy, GaussianFlowNLL
    z, cfg, model = _model_small()
    model.mode = 'training'
    model.train()
    eps = torch.from_numpy(z['in/eps']).float()
    model.reparameterize = lambda mu, logvar: eps * torch.exp(0.5 * logvar) + mu
    g_in = torch.from_numpy(z['in/g_input']).float()
    enc = model.encode(g_in)
    assert max_rel(enc['g_posterior_samples'].detach(), torch.from_numpy(z['train/g_sample']), floor=1e-3) < 2e-5
    assert max_rel(enc['g_prior_samples'][0].detach(), torch.from_numpy(z['train/g_prior_z']), floor=1e-3) < 5e-5
    gnll = GaussianFlowNLL()(enc['g_prior_samples'], enc['g_prior_mus'], enc['g_prior_logvars'])
    gent = GaussianEntropy()(enc['g_posterior_logvars'])
    assert abs(float(gnll) - float(z['train/gnll'])) < 1e-4 * abs(float(z['train/gnll']))
    assert abs(float(gent) - float(z['train/gent'])) < 1e-4 * abs(float(z['train/gent']))
    logits = model.get_weights(enc['g_posterior_samples'])
    assert max_rel(logits.detach(), torch.from_numpy(z['train/logits']), floor=1e-3) < 5e-5
    logits = model.get_weights(enc['g_posterior_samples'], warmup=True)
    assert max_rel(logits.detach(), torch.from_numpy(z['warmup/logits']), floor=1e-3) < 1e-6


def test_flow_mixture_loss_combines_terms_like_the_reference():
    """losses.py:159-173 on list-style decoder outputs built from the golden's decode-level case (no kernel)."""
    from go_with_the_flows_b200.networks.losses import Flow_Mixture_Loss
    from oracle import flow_oracle as fo
    from tests.util import Golden
    gd = Golden('small_free_learned')
    sd = gd.sd()
    p, g = gd.t('in/p'), gd.t('in/g')
    out = fo.mixture_nll(p, g, sd, base_type='free', weights_type='learned_weights', training=False)
    K = gd.meta['n_components']
    dims = fo.infer_dims(sd)
    mu_b, lv_b = fo.base_gaussian(g, sd, 'free', False, None, None)
    B, _, N = p.shape
    dec = []
    for j in range(K):
        _, zj, Sj = fo.component_logp(p, g, sd, j, dims.n_flows, mu_b, lv_b, False)
        dec.append({'p_prior_samples': [zj, p], 'p_prior_mus': [mu_b.unsqueeze(2).expand(B, 3, N)],
                    'p_prior_logvars': [lv_b.unsqueeze(2).expand(B, 3, N), Sj - lv_b.unsqueeze(2)]})
    prior = {'g_prior_samples': [g], 'g_prior_mus': [torch.zeros_like(g)], 'g_prior_logvars': [torch.zeros_like(g)],
             'g_posterior_logvars': torch.zeros_like(g)}
    loss, pnll, gnll, gent = Flow_Mixture_Loss(pnll_weight=2.0, gnll_weight=0.5, gent_weight=0.25, n_components=K)(
        prior, dec, out['logits'])
    assert abs(float(pnll) - float(gd.t('eval/pnll'))) < 1e-9 * abs(float(gd.t('eval/pnll')))
    assert abs(float(loss) - (2.0 * float(pnll) + 0.5 * float(gnll) - 0.25 * float(gent))) < 1e-9 * abs(float(loss))


# ------------------------------------------------------------------------------------------ SVR image branch
def _svr_model():
    from go_with_the_flows_b200 import configs
    from go_with_the_flows_b200.networks.flow_mixture import Flow_Mixture_SVR_Model
    cfg = dict(configs.SVR)
    torch.manual_seed(0)
    return cfg, Flow_Mixture_SVR_Model(**cfg)


def test_svr_image_branch_matches_reference_latent():
    """flow_mixture.py:198-230 in 'reconstruction' mode: image -> ResNet18 -> g0_prior -> prior flow (direct) ->
    latent and mixture logits, on weights from torch.manual_seed(0) + the constructor (config_SVR.yaml size)."""
    z = _npz('svr_full')
    cfg, model = _svr_model()
    assert sum(p.numel() for p in model.parameters()) == 25979835
    model.mode = 'reconstruction'
    model.eval()
    gen = torch.Generator().manual_seed(int(z['in/image_seed']))
    images = torch.randn(2, 4, 224, 224, generator=gen)
    with torch.no_grad():
        enc = model.encode(None, images)
        g = enc['g_prior_samples'][-1]
        logits = model.get_weights(g)
    assert max_rel(enc['g_prior_mus'][0], torch.from_numpy(z['svr/img_features_mu0']), floor=1e-3) < 1e-4
    assert max_rel(g, torch.from_numpy(z['svr/g']), floor=1e-3) < 1e-4
    assert max_rel(logits, torch.from_numpy(z['svr/logits']), floor=1e-3) < 1e-4


@pytest.mark.skipif(not HAVE_REF, reason='needs /root/reference (build container only)')
def test_svr_and_autoencoding_seeded_init_is_bit_identical_to_the_reference():
    import yaml
    _, fm, _, _ = _ref_modules()
    from go_with_the_flows_b200 import configs
    from go_with_the_flows_b200.networks.flow_mixture import Flow_Mixture_Model, Flow_Mixture_SVR_Model
    for yml, ours_cfg, ref_cls, our_cls in (('config_SVR.yaml', configs.SVR, fm.Flow_Mixture_SVR_Model, Flow_Mixture_SVR_Model),
                                            ('config_autoencoding.yaml', configs.AUTOENCODING, fm.Flow_Mixture_Model,
                                             Flow_Mixture_Model)):
        cfg = yaml.safe_load(open(os.path.join(REF, 'configs', yml)))
        cfg['weights_type'] = 'learned_weights'
        cfg['util_mode'] = 'training'
        # the restated config carries every value the constructors read
        for k, v in ours_cfg.items():
            if k in cfg and k not in ('util_mode',):
                assert cfg[k] == v, (yml, k, cfg[k], v)
        torch.manual_seed(0)
        ref = ref_cls(**cfg).state_dict()
        torch.manual_seed(0)
        own = our_cls(**dict(ours_cfg)).state_dict()
        assert list(ref.keys()) == list(own.keys())
        for k in ref:
            assert torch.equal(ref[k], own[k]), k


@pytest.mark.skipif(not HAVE_REF, reason='needs /root/reference (build container only)')
def test_pointnet_and_feature_encoders_match_reference_modules():
    renc, _, rlosses, _ = _ref_modules()
    from go_with_the_flows_b200.networks import encoders, losses
    torch.manual_seed(5)
    a = renc.PointNetCloudEncoder(3, 8, [16, 32])
    torch.manual_seed(5)
    b = encoders.PointNetCloudEncoder(3, 8, [16, 32])
    b.load_state_dict(a.state_dict(), strict=True)
    x = torch.randn(4, 3, 77)
    for mode in (True, False):
        a.train(mode)
        b.train(mode)
        assert torch.allclose(a(x.clone()), b(x.clone()), atol=1e-6)
    torch.manual_seed(6)
    fa = renc.WeightsEncoder(3, 16, 4, deterministic=True, mu_weight_std=0.001, mu_bias=0.0)
    torch.manual_seed(6)
    fb = encoders.WeightsEncoder(3, 16, 4, deterministic=True, mu_weight_std=0.001, mu_bias=0.0)
    for k, v in fa.state_dict().items():
        assert torch.equal(v, fb.state_dict()[k]), k
    h = torch.randn(5, 16)
    assert torch.allclose(fa(h), fb(h), atol=1e-6)
    # latent-side losses
    s = [torch.randn(5, 16) for _ in range(3)]
    m = [torch.randn(5, 16) for _ in range(3)]
    lv = [0.3 * torch.randn(5, 16) for _ in range(3)]
    assert torch.allclose(rlosses.GaussianFlowNLL()(s, m, lv), losses.GaussianFlowNLL()(s, m, lv), rtol=1e-6)
    assert torch.allclose(rlosses.GaussianEntropy()(lv[0]), losses.GaussianEntropy()(lv[0]), rtol=1e-6)
