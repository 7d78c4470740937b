"""CPU: drop-in surface (names, shapes, constructors, loader, C-ABI exports) -- no kernel runs."""
import ctypes
import os
import re

import pytest
import torch

from tests.util import GOLDEN_CASES, Golden, build_dropin, golden_cfg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('case', GOLDEN_CASES)
def test_state_dict_is_reference_compatible(case):
    gd = Golden(case)
    model = build_dropin(gd)           # strict=True load of a reference-generated state_dict
    ref_sd = gd.sd(torch.float32)
    own = model.state_dict()
    assert list(own.keys()) == list(ref_sd.keys())
    for k in own:
        assert own[k].shape == ref_sd[k].shape, k


def test_derived_decoder_size_matches_reference_configs():
    from go_with_the_flows_b200.networks.flow_mixture import Flow_Mixture_Model
    # airplane -> (11 triples, F=37), autoencoding / SVR -> (11, 33)
    probe = Flow_Mixture_Model.__new__(Flow_Mixture_Model)
    for G, want in ((128, (11, 37)), (512, (11, 33))):
        probe.n_components, probe.params_reduce_mode = 4, 'depth_and_feature'
        probe.p_decoder_n_flows, probe.p_decoder_n_features, probe.g_latent_space_size = 21, 64, G
        assert probe._get_decoder_params() == want


def test_cabi_exports_every_declared_symbol():
    from go_with_the_flows_b200 import _native
    from go_with_the_flows_b200.build import build
    build()
    header = open(os.path.join(ROOT, 'include', 'gwtf.h')).read()
    declared = set(re.findall(r'\b(gwtf_[a-z0-9_]+)\s*\(', header))
    declared -= {'gwtf_stack_desc'}
    handle = ctypes.CDLL(_native.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), name
    assert set(_native.EXPORTED) == declared
    assert _native.lib().gwtf_version() == 2


def test_flow_modules_refuse_cpu_tensors():
    from go_with_the_flows_b200._native import GwtfError
    gd = Golden(GOLDEN_CASES[0])
    model = build_dropin(gd)
    p = gd.t('in/p', torch.float32)
    g = gd.t('in/g', torch.float32)
    with pytest.raises(GwtfError):
        model.decode(p, g, p.shape[2])


@pytest.mark.skipif(not os.path.isdir('/root/reference/lib/networks'), reason='reference tree only in the build container')
def test_seeded_construction_reproduces_reference_weights():
    import sys
    sys.dont_write_bytecode = True
    sys.path.insert(0, '/root/reference')
    try:
        from lib.networks.flow_mixture import Flow_Mixture_Model as RefModel
        from go_with_the_flows_b200.networks.flow_mixture import Flow_Mixture_Model
        cfg = golden_cfg(Golden('small_freevar_global'))
        torch.manual_seed(3)
        ref = RefModel(**cfg)
        torch.manual_seed(3)
        own = Flow_Mixture_Model(**cfg)
        rs, os_ = ref.state_dict(), own.state_dict()
        assert list(rs.keys()) == list(os_.keys())
        for k in rs:
            assert torch.equal(rs[k], os_[k]), k
    finally:
        sys.path.remove('/root/reference')
        for m in [m for m in sys.modules if m == 'lib' or m.startswith('lib.')]:
            del sys.modules[m]


def test_flat_storage_survives_syncbn_conversion_and_zero_grad():
    """The decoder tensors live in flat masters (views keep the module API).  Converting to
    SyncBatchNorm AFTER the stack exists (train_ae.py:152 does it right before DDP), moving values
    with load_state_dict, and zero_grad(set_to_none=True) must all keep masters and modules in sync."""
    gd = Golden('small_free_learned')
    model = build_dropin(gd)
    stack = model.flow_stack()
    stack.prepare()
    ref = {k: v.clone() for k, v in model.state_dict().items()}
    model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    stack.prepare()
    assert stack._is_flat()
    for k, v in model.state_dict().items():
        assert torch.equal(v, ref[k]), k
    # masters and module tensors are the same memory
    key = 'pc_decoder.1.flows.1.nvp3.T_logvar_0.logvar_sd1_bn.running_var'
    stack.masters['bn'].tensor.mul_(2.0)
    assert torch.allclose(model.state_dict()[key], 2.0 * ref[key])
    model.load_state_dict(ref, strict=True)
    assert stack._is_flat()
    assert torch.allclose(model.state_dict()[key], ref[key])
    # gradients: autograd on the masters, .grad views on the parameters, reset by zero_grad
    g = gd.t('in/g', torch.float32)
    model.eval()
    stack.film(g, False, False).sum().backward()
    w = model.pc_decoder[2].flows[0].nvp2.T_mu_0_cond_b[3].weight
    g1 = w.grad.clone()
    assert float(g1.abs().sum()) > 0
    model.zero_grad(set_to_none=True)
    stack.prepare()
    stack.film(g, False, False).sum().backward()
    assert torch.allclose(w.grad, g1)            # not accumulated on top of the stale master gradient
    stack.film(g, False, False).sum().backward()
    assert torch.allclose(w.grad, 2 * g1)        # but accumulated across backward calls like autograd does


def test_cabi_host_side_queries_without_a_gpu():
    """Entry points that only do host arithmetic / argument checks: engine resolution, kept-activation
    buffer and workspace sizes, exchange handle validation (no kernel is launched)."""
    import ctypes
    from go_with_the_flows_b200 import _native as nat
    lib = nat.lib()
    desc = nat.StackDesc()
    desc.n_components, desc.n_layers, desc.n_features = 4, 33, 37
    desc.rec_stride = lib.gwtf_rec_stride(37)
    for l in range(33):
        desc.warp_mask[l] = 1 << (l % 3)
    d = ctypes.byref(desc)
    desc.engine = nat.ENGINE_TC_FWD
    assert lib.gwtf_resolved_engine(d, 0) == nat.ENGINE_TC_FWD and lib.gwtf_resolved_engine(d, 1) == nat.ENGINE_MMA
    # L * K * 2 nets * B * roundup256(N) * roundup8(F) floats under the mma.sync backward
    assert lib.gwtf_keep_floats(d, 64, 2048) == 33 * 4 * 2 * 64 * 2048 * 40
    assert lib.gwtf_keep_floats(d, 3, 50) == 33 * 4 * 2 * 3 * 256 * 40
    desc.engine = nat.ENGINE_FMA                  # FMA engine: (L,K,2,F,B,N)
    assert lib.gwtf_keep_floats(d, 3, 50) == 33 * 4 * 2 * 37 * 3 * 50
    desc.engine = nat.ENGINE_DEFAULT              # tcgen05 forward and backward: the backward always recomputes
    assert lib.gwtf_resolved_engine(d, 0) == nat.ENGINE_TC_FWD and lib.gwtf_resolved_engine(d, 1) == nat.ENGINE_TC
    assert lib.gwtf_keep_floats(d, 3, 50) == 0
    desc.n_features, desc.rec_stride = 44, lib.gwtf_rec_stride(44)      # wider than the tcgen05 tile: mma.sync
    assert lib.gwtf_resolved_engine(d, 0) == nat.ENGINE_MMA and lib.gwtf_resolved_engine(d, 1) == nat.ENGINE_MMA
    desc.n_features, desc.rec_stride = 37, lib.gwtf_rec_stride(37)
    desc.engine = 1                               # the retired one-tile-per-CTA engine is not a valid choice
    assert lib.gwtf_resolved_engine(d, 0) == -1
    desc.engine = nat.ENGINE_DEFAULT
    assert lib.gwtf_eval_layers_workspace_bytes(d, 4, 2048) == 4 * (2 * 4 * 4 * 3 * 2048 + 4 * 4 * 2048)
    npad = 2560 + 128 * 4                          # roundup128(2500) + 128 K
    assert lib.gwtf_sample_workspace_bytes(d, 2, 2500) >= 4 * (2 * 2 * 3 * npad + 2 * 2500)
    out = ctypes.c_void_p()
    assert lib.gwtf_exchange_world(None) == 1
    assert lib.gwtf_exchange_create(0, 1, None, None, 0, 0.0, ctypes.byref(out)) == 0   # single rank: no peers needed
    assert lib.gwtf_exchange_world(out) == 1 and lib.gwtf_exchange_seq(out) == 0
    assert lib.gwtf_exchange_resync(out, 5) == 0 and lib.gwtf_exchange_seq(out) == 5
    assert lib.gwtf_exchange_resync(out, 3) != 0                                         # never backwards
    assert lib.gwtf_exchange_destroy(out) == 0
    assert lib.gwtf_exchange_create(0, 4, None, None, 0, 0.0, ctypes.byref(out)) != 0   # several ranks need peer buffers
    assert lib.gwtf_exchange_create(5, 4, None, None, 0, 0.0, ctypes.byref(out)) != 0
    assert b'rank' in lib.gwtf_last_error_string()


@pytest.mark.parametrize('slots', [2, 4])
def test_tile_schedule_of_the_persistent_kernels(slots):
    """Host replay of the kernels' own tile scheduling code (gwtf_debug_tile_schedule): over launch shapes that give
    CTAs odd tile counts, tile ranges starting mid-shape and shapes with odd tile counts, every tile is processed
    exactly once, by the CTA that owns its range; and in the two-slot backward kernel consecutive point
    contractions of a CTA always belong to different slots (they take turns on one operand buffer by mbarrier
    phase parity -- two in a row in one slot is the deadlock of 16 clouds x 2048 points on 148 SMs)."""
    import ctypes
    import numpy as np
    from go_with_the_flows_b200 import _native as nat
    lib = nat.lib()
    rng = np.random.RandomState(7 + slots)
    cases = [(16, 16, 37), (24, 16, 37), (5, 16, 37), (64, 16, 37), (7, 16, 49), (2, 9, 9), (13, 20, 37), (3, 2, 3),
             (9, 1, 5), (40, 16, 37), (1, 16, 8), (11, 3, 37)]
    cases += [(int(rng.randint(1, 70)), int(rng.randint(1, 24)), int(rng.randint(1, 75))) for _ in range(60)]
    for B, tps, gx_max in cases:
        total = B * tps
        gx = max(1, min(gx_max, (total + slots - 1) // slots))           # the launchers' grid rule
        cta = np.full(total, -1, np.int32)
        slot = np.full(total, -1, np.int32)
        seq = np.full(total, -1, np.int32)
        ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)                # noqa: E731
        n = lib.gwtf_debug_tile_schedule(total, tps, gx, slots, ptr(cta), ptr(slot), ptr(seq))
        assert n == total, (B, tps, gx, n, lib.gwtf_last_error_string())
        assert (cta >= 0).all() and (slot >= 0).all() and (slot < slots).all()
        per = (total + gx - 1) // gx
        assert (cta == np.arange(total) // per).all()                    # contiguous ranges
        for c in range(gx):
            idx = np.nonzero(cta == c)[0]
            if slots == 2 and len(idx):
                order = idx[np.argsort(seq[idx])]
                assert (np.sort(seq[idx]) == np.arange(len(idx))).all(), (B, tps, gx, c)
                assert (slot[order] == np.arange(len(idx)) % 2).all(), (B, tps, gx, c, slot[order])
                assert (order == idx).all()                              # contractions run in tile order
