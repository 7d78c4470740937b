"""CPU: the oracle restatement replays the golden vectors produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from oracle import flow_oracle as fo
from tests.util import GOLDEN_CASES, Golden, max_rel, rel_l2


def _run(gd, training, dtype=torch.float64, warmup=False):
    sd = gd.sd(dtype)
    for k, v in sd.items():
        if v.is_floating_point() and k.startswith(('pc_decoder', 'p_prior', 'mixture_weights')) \
                and 'running' not in k and not k.endswith('eps'):
            v.requires_grad_(True)
    p = gd.t('in/p', dtype).requires_grad_(True)
    g = gd.t('in/g', dtype).requires_grad_(True)
    upd = fo.BNUpdates()
    out = fo.mixture_nll(p, g, sd, base_type=gd.meta['p_decoder_base_type'],
                         weights_type=gd.meta['weights_type'], warmup=warmup, training=training,
                         upd=upd, base_var=gd.meta['p_decoder_base_var'])
    return sd, p, g, out, upd


@pytest.mark.parametrize('case', GOLDEN_CASES)
@pytest.mark.parametrize('tag', ['train', 'eval'])
def test_nll_forward_backward_fp64(case, tag):
    gd = Golden(case)
    sd, p, g, out, upd = _run(gd, tag == 'train')
    assert max_rel(out['logp'], gd.t(f'{tag}/logp')) < 1e-9
    assert max_rel(out['nll'], gd.t(f'{tag}/nll')) < 1e-9
    assert abs(float(out['pnll']) - float(gd.t(f'{tag}/pnll'))) < 1e-9 * abs(float(out['pnll']))
    out['pnll'].backward()
    assert rel_l2(p.grad, gd.t(f'{tag}/dp')) < 1e-8
    assert rel_l2(g.grad, gd.t(f'{tag}/dg')) < 1e-8
    for k in gd.keys(f'{tag}/grad/'):
        ref = gd.t(f'{tag}/grad/{k}')
        got = sd[k].grad if sd[k].grad is not None else torch.zeros_like(ref)
        if float(ref.norm()) == 0:
            assert float(got.norm()) < 1e-12, k
        else:
            assert rel_l2(got, ref) < 1e-7, k


@pytest.mark.parametrize('case', GOLDEN_CASES)
def test_bn_running_stat_updates(case):
    gd = Golden(case)
    sd, p, g, out, upd = _run(gd, True)
    keys = [k for k in gd.keys('train/bn/') if k.startswith(('pc_decoder', 'p_prior', 'mixture_weights'))]
    assert keys
    for k in keys:
        ref = gd.t(f'train/bn/{k}')
        if k.endswith('num_batches_tracked'):
            assert int(sd[k]) + int(upd[k]) == int(ref), k
        else:
            assert max_rel(upd[k], ref, floor=1e-6) < 1e-9, k


@pytest.mark.parametrize('case', GOLDEN_CASES)
def test_warmup_uses_global_logits(case):
    gd = Golden(case)
    _, _, _, out, _ = _run(gd, False, warmup=True)
    assert abs(float(out['pnll']) - float(gd.t('eval_warmup/pnll'))) < 1e-9 * abs(float(out['pnll']))


@pytest.mark.parametrize('case', GOLDEN_CASES)
def test_fp32_oracle_within_noise_floor(case):
    gd = Golden(case)
    _, _, _, out, _ = _run(gd, True, torch.float32)
    assert max_rel(out["nll"].detach(), gd.t("train/nll")) < 2e-4


@pytest.mark.parametrize('case', GOLDEN_CASES)
def test_sampling_matches_reference_scatter(case):
    gd = Golden(case)
    sd = gd.sd()
    x, labels, _ = fo.sample(gd.t('in/g'), sd, gd.z['sample/idx'], gd.t('sample/eps'),
                             base_type=gd.meta['p_decoder_base_type'], base_var=gd.meta['p_decoder_base_var'])
    assert max_rel(x, gd.t('sample/x'), floor=1e-6) < 1e-9
    assert torch.equal(labels, gd.t('sample/labels'))


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        got = fo.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert tuple(int(x) for x in got) == want


def test_component_index_is_numpy_choice_rule():
    rs = np.random.RandomState(0)
    logits = rs.randn(5).astype(np.float32)
    cdf = fo.mixture_cdf(logits)
    u = rs.rand(10000).astype(np.float32)
    e = np.exp(logits)
    probs = (e / e.sum()).astype(np.float64)
    want = np.cumsum(probs).searchsorted(u, side='right')
    got = fo.component_index(cdf, u)
    # identical except where float32 rounding of the cdf moves a boundary across a sample
    assert (got != np.minimum(want, 4)).mean() < 1e-3
    assert got.min() >= 0 and got.max() <= 4
    u_edge = np.array([0.0, np.float32(1.0) - np.float32(2.0 ** -24)], dtype=np.float32)
    assert fo.component_index(cdf, u_edge).tolist() == [0, 4]


def test_box_muller_moments():
    _, w = fo.sample_streams(2026, 0, 4, 4096)
    e = fo.box_muller(w)
    assert e.shape == (4, 3, 4096)
    assert abs(e.mean()) < 0.02 and abs(e.std() - 1.0) < 0.02
