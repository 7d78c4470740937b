"""GPU parity through the entry points the reference's CLIs actually call, against golden vectors
generated from the unmodified reference (tests/golden/make_golden_r2.py) and the fp64 CPU oracle:

  model(g_input, p_input, ...) + Flow_Mixture_Loss + backward (models.py:224-258, losses.py:159-173, training.py:40-54)
  one_flow_decode + PointFlowNLL under autograd (models.py:153-207, losses.py:7-20)
  Flow_Mixture_SVR_Model sampling at config_SVR.yaml size, 2500 points (flow_mixture.py:141-177,198-230)
  config_autoencoding.yaml (C3: F=33, G=512, freevar) full size against the fp64 oracle
  gradients at ragged cloud sizes, the fused AMSGrad step, seeded sampling streams, the non-finite counter."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import flow_oracle as fo
from tests import parity
from tests.util import GOLDEN_DIR, Golden, build_dropin, max_rel, nll_err, rel_l2

pytestmark = pytest.mark.gpu

ENGINES = ['tcgen05', 'tcgen05_fwd', 'mma', 'fma']


@pytest.fixture(params=ENGINES, autouse=True)
def contraction_engine(request):
    from go_with_the_flows_b200 import _native as nat
    from go_with_the_flows_b200 import flowstack
    prev = flowstack.set_default('engine', {'tcgen05': nat.ENGINE_TC, 'tcgen05_fwd': nat.ENGINE_TC_FWD,
                                            'mma': nat.ENGINE_MMA, 'fma': nat.ENGINE_FMA}[request.param])
    yield request.param
    flowstack.set_default('engine', prev)


def _npz(name):
    return np.load(os.path.join(GOLDEN_DIR, name + '.npz'))


def _group_err(named, z, prefix, keys):
    num = den = 0.0
    for k in keys:
        ref = torch.from_numpy(z[prefix + k]).double()
        got = named[k].grad
        got = torch.zeros_like(ref) if got is None else got.detach().cpu().double()
        num += float((got - ref).pow(2).sum())
        den += float(ref.pow(2).sum())
    return (num / max(den, 1e-300)) ** 0.5


# ------------------------------------------------------------------------------------------ model() + loss
@pytest.mark.parametrize('tag', ['train', 'warmup'])
def test_model_call_loss_and_backward_match_the_reference(tag):
    """The call training.py:40-54 makes: model(g_clouds, p_clouds, ..., warmup) -> Flow_Mixture_Loss -> backward."""
    from go_with_the_flows_b200.networks.flow_mixture import Flow_Mixture_Model
    from go_with_the_flows_b200.networks.losses import Flow_Mixture_Loss
    z = _npz('model_small')
    cfg = ast.literal_eval(str(z['meta']))
    model = Flow_Mixture_Model(**cfg)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd/')}
    model.load_state_dict({k: (v.float() if v.is_floating_point() else v) for k, v in sd.items()}, strict=True)
    model = model.cuda().train()
    model.mode = 'training'
    eps = torch.from_numpy(z['in/eps']).float().cuda()
    model.reparameterize = lambda mu, logvar: eps * torch.exp(0.5 * logvar) + mu
    g_in = torch.from_numpy(z['in/g_input']).float().cuda()
    p_in = torch.from_numpy(z['in/p_input']).float().cuda()
    out_enc, out_dec, logits = model(g_in, p_in, images=None, n_sampled_points=None, labeled_samples=False,
                                     warmup=(tag == 'warmup'))
    assert isinstance(out_dec, list) and len(out_dec) == cfg['n_components']
    loss, pnll, gnll, gent = Flow_Mixture_Loss(**cfg)(out_enc, out_dec, logits)
    assert not torch.isnan(loss)
    loss.backward()
    for name, val in (('loss', loss), ('pnll', pnll), ('gnll', gnll), ('gent', gent)):
        want = float(z[f'{tag}/{name}'])
        assert abs(float(val) - want) < 2e-5 * max(1.0, abs(want)), (name, float(val), want)
    named = dict(model.named_parameters())
    keys = [k[len(tag) + 6:] for k in z.files if k.startswith(f'{tag}/grad/')]
    groups = {'decoder': [k for k in keys if k.startswith('pc_decoder')],
              'encoder': [k for k in keys if k.startswith(('pc_encoder', 'g_posterior'))],
              'prior': [k for k in keys if k.startswith(('g_prior', 'g0_prior'))],
              'heads': [k for k in keys if k.startswith(('p_prior', 'mixture_weights'))]}
    assert sum(len(v) for v in groups.values()) == len(keys)
    for gname, gkeys in groups.items():
        err = _group_err(named, z, f'{tag}/grad/', gkeys)
        assert err < 2e-4, (gname, err)
    sd_after = model.state_dict()
    for k in [k[len(tag) + 4:] for k in z.files if k.startswith(f'{tag}/bn/')]:
        ref = torch.from_numpy(z[f'{tag}/bn/{k}'])
        if k.endswith('num_batches_tracked'):
            assert int(sd_after[k]) == int(ref), k
        else:
            assert max_rel(sd_after[k].cpu(), ref, floor=1e-3) < 1e-4, k
    assert model.nonfinite_points() == 0


# ------------------------------------------------------------------------------------------ one_flow_decode
@pytest.mark.parametrize('tag', ['train', 'eval'])
def test_one_flow_decode_is_differentiable_like_the_reference(tag):
    """models.py:153-207 with ONE decoder + PointFlowNLL (losses.py:7-20): values and every gradient."""
    from go_with_the_flows_b200.networks.losses import PointFlowNLL
    z = _npz('oneflow_small')
    gd = Golden('small_free_learned')
    model = build_dropin(gd, 'cuda')
    model.mode = 'training'
    model.train(tag == 'train')
    p = torch.from_numpy(z['in/p']).float().cuda().requires_grad_(True)
    g = torch.from_numpy(z['in/g']).float().cuda().requires_grad_(True)
    out = model.one_flow_decode(p, g, model.pc_decoder[1], p.shape[2])
    assert len(out['p_prior_samples']) == int(z[f'{tag}/n_entries'])
    assert len(out['p_prior_mus']) == len(out['p_prior_logvars']) == int(z[f'{tag}/n_entries'])
    nll = PointFlowNLL()(out)
    nll.sum(dim=2).mean().backward()
    assert nll_err(nll.detach().cpu(), torch.from_numpy(z[f'{tag}/nll'])) < 1e-4
    zref = torch.from_numpy(z[f'{tag}/z'])
    assert float((out['p_prior_samples'][0].detach().cpu().double() - zref).abs().max() / zref.abs().max()) < 1e-4
    assert rel_l2(p.grad.cpu(), torch.from_numpy(z[f'{tag}/dp'])) < 2e-4
    assert rel_l2(g.grad.cpu(), torch.from_numpy(z[f'{tag}/dg'])) < 2e-4
    named = dict(model.named_parameters())
    keys = [k[len(tag) + 6:] for k in z.files if k.startswith(f'{tag}/grad/')]
    dec = [k for k in keys if k.startswith('pc_decoder.1.')]
    assert dec and _group_err(named, z, f'{tag}/grad/', dec) < 2e-4
    other = [k for k in keys if k.startswith('p_prior')]
    assert _group_err(named, z, f'{tag}/grad/', other) < 5e-4
    # the components that did not take part got no gradient
    assert all(named[k].grad is None or float(named[k].grad.abs().sum()) == 0.0
               for k in named if k.startswith('pc_decoder.0.'))


def test_direct_mode_in_training_refuses_to_drop_gradients():
    from go_with_the_flows_b200 import _native as nat
    gd = Golden('small_free_learned')
    model = build_dropin(gd, 'cuda').train()
    p = gd.t('in/p', torch.float32, 'cuda')
    g = gd.t('in/g', torch.float32, 'cuda')
    with pytest.raises(nat.GwtfError):
        model.pc_decoder[0](p, g, mode='direct')
    with torch.no_grad():
        ps, _, _ = model.pc_decoder[0](p, g, mode='direct')
    assert torch.isfinite(ps[-1]).all()


# ------------------------------------------------------------------------------------------ SVR (C4)
def test_svr_model_samples_2500_points_per_shape():
    """config_SVR.yaml size: K=4, L=33, F=33, G=512, freevar base, cloud_size 2500 (not a tile multiple)."""
    from go_with_the_flows_b200 import configs
    from go_with_the_flows_b200.flowstack import sample_mixture
    from go_with_the_flows_b200.networks.flow_mixture import Flow_Mixture_SVR_Model
    z = _npz('svr_full')
    cfg = dict(configs.SVR)
    torch.manual_seed(0)
    model = Flow_Mixture_SVR_Model(**cfg).cuda().eval()
    model.mode = 'reconstruction'
    gen = torch.Generator().manual_seed(int(z['in/image_seed']))
    images = torch.randn(2, 4, 224, 224, generator=gen).cuda()
    N = cfg['cloud_size']
    with torch.no_grad():
        enc = model.encode(None, images)
        g = enc['g_prior_samples'][-1]
        assert max_rel(g.cpu(), torch.from_numpy(z["svr/g"]), floor=1e-2) < 1e-2       # cuDNN convolutions vs CPU
        # decode on the reference's latent so the comparison isolates the flow stack
        g = torch.from_numpy(z['svr/g']).cuda()
        logits = model.get_weights(g)
        mu_b, lv_b = model.base_gaussian(g)
        idx = torch.from_numpy(z['sample/idx'])
        eps = torch.from_numpy(z['sample/eps'])
        x, labels, _ = sample_mixture(model.flow_stack(), g, mu_b, lv_b, logits, N, seed=1, idx=idx, eps=eps)
    ref = torch.from_numpy(z['sample/x'])
    assert float((x.cpu().double() - ref).abs().max() / ref.abs().max()) < 1e-4
    assert torch.equal(labels.cpu().double(), torch.from_numpy(z['sample/labels']).double())
    # and the call evaluate / reconstruct make: model(..., images, n_sampled_points, labeled_samples=True)
    model.sample_seed = 2026
    with torch.no_grad():
        out_enc, samples, lab, lg = model(None, torch.zeros(2, 3, N, device='cuda'), images=images, n_sampled_points=N,
                                          labeled_samples=True)
    assert samples.shape == (2, 3, N) and lab.shape == (2, N) and lab.dtype == samples.dtype
    assert torch.isfinite(samples).all() and int(lab.min()) >= 1 and int(lab.max()) <= cfg['n_components']
    # labels are bit-exact against the CPU Philox restatement fed with the same logits
    u, _ = fo.sample_streams(2026, 0, 2, N)
    host_logits = lg.float().cpu().numpy()
    want = np.stack([fo.component_index(fo.mixture_cdf(host_logits[b]), u[b]) for b in range(2)]) + 1
    assert np.array_equal(lab.cpu().numpy().astype(np.int64), want.astype(np.int64))


# ------------------------------------------------------------------------------------------ C3 full size
def test_c3_autoencoding_4x2048_against_fp64_oracle():
    """config_autoencoding.yaml model (K=4, L=33, F=33, G=512, freevar), random init, train mode."""
    from go_with_the_flows_b200 import configs
    from go_with_the_flows_b200.networks.flow_mixture import Flow_Mixture_Model
    from go_with_the_flows_b200.networks.losses import FlowMixtureNLL
    from tests.test_gpu_fullsize import _grad_err, oracle_case
    cfg, p, g, (want, dp64, dg64, sd64), (w32, dp32, dg32, sd32) = oracle_case('autoencoding', 4, 2048, True)
    torch.manual_seed(0)
    model = Flow_Mixture_Model(**dict(configs.AUTOENCODING))
    assert model.flow_stack().F == 33 and model.flow_stack().L == 33 and model.flow_stack().G == 512
    model = model.cuda().train()
    model.mode = 'training'
    pc = p.cuda().requires_grad_(True)
    gc = g.cuda().requires_grad_(True)
    out, logits = model.decode(pc, gc, 2048)
    FlowMixtureNLL()(out, logits).backward()
    assert nll_err(out[0]['mixture_nll'].detach().cpu(), want['nll']) < 1e-4
    assert rel_l2(pc.grad.cpu(), dp64) < max(1e-4, 3 * rel_l2(dp32, dp64))
    assert rel_l2(gc.grad.cpu(), dg64) < max(1e-4, 3 * rel_l2(dg32, dg64))
    named = {k: v.grad.detach().cpu() for k, v in model.named_parameters() if v.grad is not None}
    is_dec = lambda k: k.startswith('pc_decoder') and k in named   # noqa: E731
    ours = _grad_err(named, sd64, is_dec)
    noise = _grad_err({k: v.grad for k, v in sd32.items() if v.requires_grad and v.grad is not None}, sd64, is_dec)
    assert ours < max(1e-4, 3 * noise), (ours, noise)


# ------------------------------------------------------------------------------------------ ragged gradients
@pytest.mark.parametrize('B,N', [(2, 1025), (2, 2500), (3, 130), (40, 2048), (7, 2048)])
def test_train_mode_gradients_at_ragged_cloud_sizes(B, N):
    """N not a multiple of 128 / 256 (config_SVR uses 2500), a batch large enough that every persistent CTA walks
    many tiles and several shapes (40 x 2048: 13 tiles per CTA), and one whose CTAs own 3 tiles each (7 x 2048 with
    K = 3 on 148 SMs), so that tile ranges start at odd offsets inside a shape and the two-slot backward kernel runs
    partial rounds: NLL and every gradient against the fp64 oracle."""
    from go_with_the_flows_b200.networks.losses import FlowMixtureNLL
    gd = Golden('small_c3_freevar')
    gen = torch.Generator().manual_seed(B * 7919 + N)
    p = 0.4 * torch.randn(B, 3, N, generator=gen, dtype=torch.float64)
    g = 0.7 * torch.randn(B, gd.meta['g_latent_space_size'], generator=gen, dtype=torch.float64)
    sd = gd.sd()
    for k, v in sd.items():
        if v.is_floating_point() and k.startswith('pc_decoder') and 'running' not in k and not k.endswith('eps'):
            v.requires_grad_(True)
    p64, g64 = p.clone().requires_grad_(True), g.clone().requires_grad_(True)
    want = fo.mixture_nll(p64, g64, sd, base_type=gd.meta['p_decoder_base_type'], weights_type=gd.meta['weights_type'],
                          training=True, base_var=gd.meta['p_decoder_base_var'])
    want['pnll'].backward()
    model = build_dropin(gd, 'cuda').train()
    model.mode = 'training'
    pc = p.float().cuda().requires_grad_(True)
    gc = g.float().cuda().requires_grad_(True)
    out, logits = model.decode(pc, gc, N)
    FlowMixtureNLL()(out, logits).backward()
    assert nll_err(out[0]['mixture_nll'].detach().cpu(), want['nll'].detach()) < 2e-4
    assert rel_l2(pc.grad.cpu(), p64.grad) < 5e-4
    assert rel_l2(gc.grad.cpu(), g64.grad) < 2e-4
    named = dict(model.named_parameters())
    num = den = 0.0
    for k, v in sd.items():
        if v.requires_grad and v.grad is not None:
            num += float((named[k].grad.detach().cpu().double() - v.grad).pow(2).sum())
            den += float(v.grad.pow(2).sum())
    assert (num / den) ** 0.5 < 2e-4


# ------------------------------------------------------------------------------------------ optimizer
def test_fused_amsgrad_over_flat_masters_equals_per_tensor_update():
    """One kernel per flat master (gwtf_adam_step) against the per-tensor multi-tensor path on a copy of the model:
    parameters and optimizer state after three steps, and the reference-layout state_dict."""
    from go_with_the_flows_b200.networks.losses import FlowMixtureNLL
    from go_with_the_flows_b200.networks.optimizers import Adam
    gd = Golden('small_free_learned')
    a, b = build_dropin(gd, 'cuda').train(), build_dropin(gd, 'cuda').train()
    a.mode = b.mode = 'training'
    p = gd.t('in/p', torch.float32, 'cuda')
    g = gd.t('in/g', torch.float32, 'cuda')
    oa = Adam(a.parameters(), lr=1e-3, betas=(0.9, 0.995), weight_decay=1e-5, amsgrad=True)
    ob = Adam(b.parameters(), lr=1e-3, betas=(0.9, 0.995), weight_decay=1e-5, amsgrad=True)
    for step in range(3):
        for model, opt in ((a, oa), (b, ob)):
            opt.zero_grad()
            out, logits = model.decode(p, g, p.shape[2])
            FlowMixtureNLL()(out, logits).backward()
        # same gradients on both sides (copy a's onto b), then fused vs per-tensor
        for (ka, pa), (kb, pb) in zip(a.named_parameters(), b.named_parameters()):
            if pa.grad is not None:
                pb.grad = pa.grad.detach().clone()      # plain tensors: b's optimizer takes the per-tensor path
        for prm in b.parameters():
            if hasattr(prm, '_gwtf_master'):
                del prm._gwtf_master
        oa.step()
        ob.step()
        assert len(oa._flat) == 6 and len(ob._flat) == 0
    for (ka, pa), (kb, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert float((pa - pb).abs().max()) <= 1e-6 * max(1.0, float(pb.abs().max())), ka
    sa, sb = oa.state_dict()['state'], ob.state_dict()['state']
    assert sa.keys() == sb.keys()
    for i in sa:
        assert int(sa[i]['step']) == int(sb[i]['step']) == 3
        for k in ('exp_avg', 'exp_avg_sq', 'max_exp_avg_sq'):
            assert sa[i][k].shape == sb[i][k].shape
            assert float((sa[i][k] - sb[i][k]).abs().max()) <= 3e-6 * max(1e-12, float(sb[i][k].abs().max())), (i, k)


# ------------------------------------------------------------------------------------------ sampling streams
def test_seeded_decode_calls_use_fresh_streams_and_replay():
    gd = Golden('small_free_learned')
    model = build_dropin(gd, 'cuda').eval()
    model.mode = 'generating'
    model.sample_seed = 1234
    g = gd.t('in/g', torch.float32, 'cuda')[:1]
    with torch.no_grad():
        a = model.decode(None, g, 500, labeled_samples=True)
        b = model.decode(None, g, 500, labeled_samples=True)
        model.sample_calls = 0
        c = model.decode(None, g, 500, labeled_samples=True)
        d = model.decode(None, g, 500, labeled_samples=True)
    assert not torch.equal(a[1], b[1]) and not torch.equal(a[0], b[0])        # consecutive calls: new draws
    assert torch.equal(a[0], c[0]) and torch.equal(a[1], c[1])                # replay from the counter
    assert torch.equal(b[0], d[0]) and torch.equal(b[1], d[1])


def test_nonfinite_counter_reports_nan_points():
    gd = Golden('small_free_learned')
    model = build_dropin(gd, 'cuda').eval()
    model.mode = 'training'
    p = gd.t('in/p', torch.float32, 'cuda').clone()
    g = gd.t('in/g', torch.float32, 'cuda')
    with torch.no_grad():
        model.decode(p, g, p.shape[2])
    assert model.nonfinite_points() == 0
    p[1, 0, 7] = float('nan')
    p[2, 2, 11] = float('inf')
    with torch.no_grad():
        out, _ = model.decode(p, g, p.shape[2])
    assert model.nonfinite_points() == 2
    assert torch.isnan(out[0]['mixture_nll'].sum())           # and the loss itself is NaN, as training.py:43 expects
    assert model.nonfinite_points() == 0
    with torch.enable_grad():
        out, _ = model.decode(p, g.clone().requires_grad_(True), p.shape[2])
    assert model.nonfinite_points() == 2
