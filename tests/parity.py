"""Shared parity measurements: our CUDA path vs the fp64 golden vectors / the CPU oracle."""
import numpy as np
import torch

from oracle import flow_oracle as fo
from tests.util import Golden, build_dropin, max_rel, nll_err, rel_l2

DECODER_PREFIXES = ('pc_decoder', 'p_prior', 'mixture_weights')


def oracle_fp32_errors(gd, tag):
    """Noise floor: the CPU oracle in fp32 against the fp64 golden (nll + gradients)."""
    training = tag == 'train'
    sd = gd.sd(torch.float32)
    for k, v in sd.items():
        if v.is_floating_point() and k.startswith(DECODER_PREFIXES) and 'running' not in k and not k.endswith('eps'):
            v.requires_grad_(True)
    p = gd.t('in/p', torch.float32).requires_grad_(True)
    g = gd.t('in/g', torch.float32).requires_grad_(True)
    out = fo.mixture_nll(p, g, sd, base_type=gd.meta['p_decoder_base_type'], weights_type=gd.meta['weights_type'],
                         training=training, base_var=gd.meta['p_decoder_base_var'])
    out['pnll'].backward()
    errs = {'nll': nll_err(out['nll'].detach(), gd.t(f'{tag}/nll')),
            'dp': rel_l2(p.grad, gd.t(f'{tag}/dp')), 'dg': rel_l2(g.grad, gd.t(f'{tag}/dg'))}
    num = den = 0.0
    for k in gd.keys(f'{tag}/grad/'):
        if not k.startswith('pc_decoder'):
            continue
        ref = gd.t(f'{tag}/grad/{k}')
        got = sd[k].grad if sd[k].grad is not None else torch.zeros_like(ref)
        num += float((got.double() - ref).pow(2).sum())
        den += float(ref.pow(2).sum())
    errs['dparams'] = (num / max(den, 1e-300)) ** 0.5
    return errs


def dropin_nll_errors(gd, tag, fused_nll=True, device='cuda'):
    """Run our drop-in model's decode + loss (+ backward) and measure against the golden."""
    from go_with_the_flows_b200.networks.losses import FlowMixtureNLL
    training = tag == 'train'
    model = build_dropin(gd, device)
    model.mode = 'training'
    model.train(training)
    model.fused_nll = fused_nll
    p = gd.t('in/p', torch.float32, device).requires_grad_(True)
    g = gd.t('in/g', torch.float32, device).requires_grad_(True)
    out_dec, logits = model.decode(p, g, p.shape[2])
    pnll = FlowMixtureNLL()(out_dec, logits)
    pnll.backward()
    res = {}
    if fused_nll:
        res['nll'] = nll_err(out_dec[0]['mixture_nll'].detach().cpu(), gd.t(f'{tag}/nll'))
    else:
        z = torch.stack([od['p_prior_samples'][0] for od in out_dec], 1)
        res['z'] = max_rel(z.detach().cpu(), gd.t(f'{tag}/z'), floor=1e-3)
    res['pnll'] = abs(float(pnll) - float(gd.t(f'{tag}/pnll'))) / abs(float(gd.t(f'{tag}/pnll')))
    res['dp'] = rel_l2(p.grad.cpu(), gd.t(f'{tag}/dp'))
    res['dg'] = rel_l2(g.grad.cpu(), gd.t(f'{tag}/dg'))
    named = dict(model.named_parameters())
    num = den = 0.0
    worst = (0.0, None)
    for k in gd.keys(f'{tag}/grad/'):
        ref = gd.t(f'{tag}/grad/{k}')
        got = named[k].grad
        got = torch.zeros_like(ref) if got is None else got.detach().cpu().double()
        e2, r2 = float((got - ref).pow(2).sum()), float(ref.pow(2).sum())
        if k.startswith('pc_decoder'):
            num += e2
            den += r2
            if r2 > 0 and (e2 / r2) ** 0.5 > worst[0]:
                worst = ((e2 / r2) ** 0.5, k)
        else:
            res.setdefault('dother', 0.0)
            if r2 > 0:
                res['dother'] = max(res['dother'], (e2 / r2) ** 0.5)
    res['dparams'] = (num / max(den, 1e-300)) ** 0.5
    res['dparams_worst'] = worst
    bn_err = 0.0
    if training:
        sd = model.state_dict()
        for k in gd.keys('train/bn/'):
            if not k.startswith(DECODER_PREFIXES):
                continue
            ref = gd.t(f'train/bn/{k}')
            if k.endswith('num_batches_tracked'):
                assert int(sd[k]) == int(ref), k
            else:
                bn_err = max(bn_err, max_rel(sd[k].cpu(), ref, floor=1e-3))
    res['bn'] = bn_err
    return res


def dropin_eval_fused_error(gd, device='cuda'):
    model = build_dropin(gd, device)
    model.mode = 'training'
    model.eval()
    p = gd.t('in/p', torch.float32, device)
    g = gd.t('in/g', torch.float32, device)
    with torch.no_grad():
        out_dec, logits = model.decode(p, g, p.shape[2])
    nll = out_dec[0]['mixture_nll']
    return nll_err(nll.cpu(), gd.t('eval/nll'))


def dropin_sample_errors(gd, device='cuda'):
    from go_with_the_flows_b200.flowstack import sample_mixture
    model = build_dropin(gd, device)
    model.mode = 'generating'
    model.eval()
    g = gd.t('in/g', torch.float32, device)
    idx = torch.from_numpy(gd.z['sample/idx'])
    eps = gd.t('sample/eps', torch.float32)
    with torch.no_grad():
        logits = model.get_weights(g)
        mu_b, lv_b = model.base_gaussian(g)
        x, labels, z = sample_mixture(model.flow_stack(), g, mu_b, lv_b, logits, eps.shape[2], seed=1, idx=idx,
                                      eps=eps, want_z=True)
    ref = gd.t('sample/x')
    err = float((x.cpu().double() - ref).abs().max() / ref.abs().max())
    lab_ok = bool(torch.equal(labels.cpu().double(), gd.t('sample/labels')))
    return err, lab_ok
