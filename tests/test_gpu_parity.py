"""GPU parity (run on the B200 box): the CUDA path behind the drop-in modules vs the fp64 golden
vectors generated from the unmodified reference, with the CPU-oracle fp32 error as the noise
yardstick: err(ours) <= max(tol, 3 * err(oracle fp32))."""
import numpy as np
import pytest
import torch

from oracle import flow_oracle as fo
from tests import parity
from tests.util import GOLDEN_CASES, Golden, build_dropin, max_rel

pytestmark = pytest.mark.gpu


@pytest.fixture(params=['tcgen05', 'tcgen05_fwd', 'mma', 'fma'], autouse=True)
def contraction_engine(request):
    """Every GPU parity test runs on every engine of the per-layer kernels: tcgen05 forward + backward (default),
    tcgen05 forward + mma.sync backward, mma.sync fragments, and the FP32 FMA pipe."""
    from go_with_the_flows_b200 import _native as nat
    from go_with_the_flows_b200 import flowstack
    prev = flowstack.set_default('engine', {'tcgen05': nat.ENGINE_TC, 'tcgen05_fwd': nat.ENGINE_TC_FWD,
                                            'mma': nat.ENGINE_MMA, 'fma': nat.ENGINE_FMA}[request.param])
    yield request.param
    flowstack.set_default('engine', prev)

NLL_TOL = 1e-4        # north star: per-point log-likelihood within 1e-4 relative (fp32 path)
GRAD_TOL = 1e-4       # norm-wise on gradients


@pytest.mark.parametrize('case', GOLDEN_CASES)
def test_fused_eval_nll(case):
    gd = Golden(case)
    noise = parity.oracle_fp32_errors(gd, 'eval')
    assert parity.dropin_eval_fused_error(gd) < max(NLL_TOL, 3 * noise['nll'])


@pytest.mark.parametrize('case', GOLDEN_CASES)
@pytest.mark.parametrize('tag', ['train', 'eval'])
@pytest.mark.parametrize('fused', [True, False])
def test_nll_forward_backward(case, tag, fused):
    gd = Golden(case)
    noise = parity.oracle_fp32_errors(gd, tag)
    res = parity.dropin_nll_errors(gd, tag, fused_nll=fused)
    if fused:
        assert res['nll'] < max(NLL_TOL, 3 * noise['nll']), res
    assert res['pnll'] < 1e-5, res
    assert res['dp'] < max(GRAD_TOL, 3 * noise['dp']), res
    assert res['dg'] < max(GRAD_TOL, 3 * noise['dg']), res
    assert res['dparams'] < max(GRAD_TOL, 3 * noise['dparams']), res
    assert res['dother'] < max(5e-4, 30 * noise['dg']), res
    assert res['bn'] < 1e-4, res


@pytest.mark.parametrize('case', GOLDEN_CASES)
def test_sampling_given_noise_and_indices(case):
    gd = Golden(case)
    err, labels_ok = parity.dropin_sample_errors(gd)
    assert err < 1e-4
    assert labels_ok


@pytest.mark.parametrize('case', GOLDEN_CASES[:2])
def test_sampling_philox_draws_are_bit_exact(case):
    """Integer component assignment under a fixed seed is bit-exact against the CPU Philox
    restatement; the in-kernel Box-Muller noise matches it to fp32 rounding."""
    from go_with_the_flows_b200.flowstack import mixture_cdf, sample_mixture
    gd = Golden(case)
    model = build_dropin(gd, 'cuda')
    model.mode = 'generating'
    model.eval()
    g = gd.t('in/g', torch.float32, 'cuda')
    B, N = g.shape[0], 1000
    seed, stream = (2026 << 32) | 12345, 3
    with torch.no_grad():
        logits = model.get_weights(g)
        mu_b, lv_b = model.base_gaussian(g)
        x, labels, z = sample_mixture(model.flow_stack(), g, mu_b, lv_b, logits, N, seed, stream, want_z=True)
    u, words = fo.sample_streams(seed, stream, B, N)
    host_logits = logits.detach().float().cpu().numpy()
    want = np.stack([fo.component_index(fo.mixture_cdf(host_logits[b]), u[b]) for b in range(B)])
    assert np.array_equal(labels.cpu().numpy(), want + 1)
    # the device-side cdf kernel and the CPU restatement agree bit for bit
    assert np.array_equal(mixture_cdf(logits).cpu().numpy(), np.stack([fo.mixture_cdf(r) for r in host_logits]))
    eps = torch.from_numpy(fo.box_muller(words)).double()
    z_want = mu_b.cpu().double().unsqueeze(2) + torch.exp(0.5 * lv_b.cpu().double()).unsqueeze(2) * eps
    assert float((z.cpu().double() - z_want).abs().max()) < 2e-5
    # and the flow applied to those draws agrees with the oracle's direct pass
    xo, _, _ = fo.sample(g.cpu().double(), gd.sd(), want, eps, base_type=gd.meta['p_decoder_base_type'],
                         base_var=gd.meta['p_decoder_base_var'])
    assert float((x.cpu().double() - xo).abs().max() / xo.abs().max()) < 1e-4


def test_sampling_same_seed_is_reproducible_and_streams_differ():
    from go_with_the_flows_b200.flowstack import sample_mixture
    gd = Golden(GOLDEN_CASES[0])
    model = build_dropin(gd, 'cuda')
    model.mode = 'generating'
    model.eval()
    g = gd.t('in/g', torch.float32, 'cuda')
    with torch.no_grad():
        logits = model.get_weights(g)
        mu_b, lv_b = model.base_gaussian(g)
        a = sample_mixture(model.flow_stack(), g, mu_b, lv_b, logits, 777, 5, 0)
        b = sample_mixture(model.flow_stack(), g, mu_b, lv_b, logits, 777, 5, 0)
        c = sample_mixture(model.flow_stack(), g, mu_b, lv_b, logits, 777, 5, 1)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert not torch.equal(a[1], c[1])


@pytest.mark.parametrize('case', GOLDEN_CASES)
def test_module_list_api_matches_oracle(case):
    """decoder(p, g, mode) keeps the reference's list contract (decoders.py:65-79)."""
    gd = Golden(case)
    model = build_dropin(gd, 'cuda').eval()
    p = gd.t('in/p', torch.float32, 'cuda')
    g = gd.t('in/g', torch.float32, 'cuda')
    sd = gd.sd()
    dims = fo.infer_dims(sd)
    for mode in ('inverse', 'direct'):
        ps, mus, lvs = model.pc_decoder[1](p, g, mode=mode)
        ops, omus, olvs = fo.decoder_stack(p.cpu().double(), g.cpu().double(), sd, 'pc_decoder.1.', dims.n_flows,
                                           mode, False)
        assert len(ps) == len(ops) == dims.n_layers
        for got, want in zip(ps + mus + lvs, ops + omus + olvs):
            assert float((got.cpu().double() - want).abs().max()) < 1e-4 * max(1.0, float(want.abs().max()))
    # single layer and triple
    layer = model.pc_decoder[0].flows[0].nvp2
    po, mu, lv = layer(p, g, mode='inverse')
    wo, wmu, wlv = fo.coupling_layer(p.cpu().double(), g.cpu().double(), sd, 'pc_decoder.0.flows.0.nvp2.',
                                     layer.warp_inds, layer.keep_inds, 'inverse', False)
    for got, want in ((po, wo), (mu, wmu), (lv, wlv)):
        assert float((got.cpu().double() - want).abs().max()) < 1e-5 * max(1.0, float(want.abs().max()))


def test_module_list_api_train_mode_updates_running_stats():
    gd = Golden('small_free_learned')
    model = build_dropin(gd, 'cuda').train()
    p = gd.t('in/p', torch.float32, 'cuda')
    g = gd.t('in/g', torch.float32, 'cuda')
    sd = gd.sd()
    upd = fo.BNUpdates()
    ops, _, _ = fo.decoder_stack(p.cpu().double(), g.cpu().double(), sd, 'pc_decoder.0.', fo.infer_dims(sd).n_flows,
                                 'inverse', True, upd)
    ps, _, _ = model.pc_decoder[0](p, g, mode='inverse')
    assert float((ps[0].cpu().double() - ops[0]).abs().max()) < 1e-4 * float(ops[0].abs().max())
    own = model.state_dict()
    for k, v in upd.items():
        if k.endswith('num_batches_tracked'):
            assert int(own[k]) == int(sd[k]) + v
        else:
            assert max_rel(own[k].cpu(), v, floor=1e-3) < 1e-4, k


def test_ragged_and_tiny_clouds():
    """N not a multiple of the tile, N = 1, B = 1 (edge cases of the tiling / masking)."""
    gd = Golden('small_free_learned')
    sd = gd.sd()
    for B, N in ((1, 1), (2, 257), (3, 1025)):
        gen = torch.Generator().manual_seed(B * 1000 + N)
        p = 0.4 * torch.randn(B, 3, N, generator=gen, dtype=torch.float64)
        g = 0.7 * torch.randn(B, 16, generator=gen, dtype=torch.float64)
        want = fo.mixture_nll(p, g, sd, base_type='free', weights_type='learned_weights', training=False)
        model = build_dropin(gd, 'cuda').eval()
        with torch.no_grad():
            out, _ = model.decode(p.float().cuda(), g.float().cuda(), N)
        assert max_rel(out[0]['mixture_nll'].cpu(), want['nll']) < 1e-4
        if B * N > 1:   # batch statistics need more than one point
            wt = fo.mixture_nll(p, g, sd, base_type='free', weights_type='learned_weights', training=True)
            model = build_dropin(gd, 'cuda').train()
            if B == 1:
                continue    # cond-net BatchNorm over a batch of one latent is undefined (torch raises too)
            out, _ = model.decode(p.float().cuda(), g.float().cuda(), N)
            assert max_rel(out[0]['mixture_nll'].detach().cpu(), wt['nll']) < 2e-4


def test_native_library_is_the_code_that_ran():
    import os
    from go_with_the_flows_b200 import _native
    maps = open('/proc/self/maps').read()
    _native.lib()
    maps = open('/proc/self/maps').read()
    assert os.path.basename(_native.LIB_PATH) in maps


def test_two_gpu_syncbn_equals_concatenated_batch():
    """Needs 2 GPUs (gpurun --gpus 2); skipped on a single-GPU box."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                          '--master-addr', '127.0.0.1', '--master-port', '29533',
                          os.path.join(root, 'tests', 'dist_gpu_parity.py')], capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.parametrize('keep', [True, False])
def test_kept_activations_path_matches(keep):
    """Both backward variants: activations kept by the forward apply pass, or recomputed (the default
    depends on the engine)."""
    gd = Golden('small_free_learned')
    noise = parity.oracle_fp32_errors(gd, 'train')
    import go_with_the_flows_b200.flowstack as fs
    orig = fs.FlowStack.__init__

    def patched(self, *a, **k):
        orig(self, *a, **k)
        self.keep_activations = keep
    fs.FlowStack.__init__ = patched
    try:
        res = parity.dropin_nll_errors(gd, 'train', fused_nll=True)
    finally:
        fs.FlowStack.__init__ = orig
    assert res['dparams'] < max(GRAD_TOL, 3 * noise['dparams']), res
    assert res['dg'] < max(GRAD_TOL, 3 * noise['dg']), res
    assert res['dp'] < max(GRAD_TOL, 3 * noise['dp']), res
