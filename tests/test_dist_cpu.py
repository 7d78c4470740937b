"""CPU, world_size 2 over gloo: host-side multi-rank logic of the flow stack (no kernels run).

R ranks with SyncBatchNorm statistics + averaged gradients must equal one process on the
concatenated batch (SURVEY.md §4 'Distributed without a cluster')."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import Golden, build_dropin


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        gd = Golden('small_freevar_global')           # B = 4 -> 2 shapes per rank
        g_all = gd.t('in/g', torch.float32)
        model = build_dropin(gd).train()
        stack = model.flow_stack()
        stack.prepare()
        g = g_all[rank * 2:(rank + 1) * 2].clone().requires_grad_(True)
        film = stack.film(g, True, True)              # SyncBN statistics over both ranks
        w = torch.linspace(0.5, 1.5, film[0].numel()).view_as(film[0])
        loss = (film * w).sum() / g.shape[0]          # local mean, as each rank's loss is
        loss.backward()                               # reduce hook averages master grads over ranks
        out = {
            'film': film.detach(), 'dg': g.grad,
            'dw0': stack.masters['c_w0'].tensor.grad.clone(), 'dbnw': stack.masters['c_bnw'].tensor.grad.clone(),
            'rm': stack.masters['c_rm'].tensor.clone(), 'rv': stack.masters['c_rv'].tensor.clone(),
            'pgrad': model.pc_decoder[1].flows[0].nvp2.T_mu_0_cond_w[0].weight.grad.clone(),
        }
        q.put((rank, out))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_two_ranks_equal_one_rank_on_concatenated_batch():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, _free_port_shared, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get() for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0

    gd = Golden('small_freevar_global')
    g = gd.t('in/g', torch.float32).requires_grad_(True)
    model = build_dropin(gd).train()
    stack = model.flow_stack()
    stack.prepare()
    film = stack.film(g, True, False)
    w = torch.linspace(0.5, 1.5, film[0].numel()).view_as(film[0])
    ((film * w).sum() / g.shape[0]).backward()
    for r in range(world):
        assert torch.allclose(got[r]['film'], film[r * 2:(r + 1) * 2].detach(), rtol=1e-4, atol=1e-5)
        # each rank's loss is its local mean: d/dg of the global-mean loss is half of it
        assert torch.allclose(got[r]['dg'] / world, g.grad[r * 2:(r + 1) * 2], rtol=1e-3, atol=1e-5)
        assert torch.allclose(got[r]['dw0'], stack.masters['c_w0'].tensor.grad, rtol=1e-3, atol=1e-5)
        assert torch.allclose(got[r]['dbnw'], stack.masters['c_bnw'].tensor.grad, rtol=1e-3, atol=1e-5)
        assert torch.allclose(got[r]['rm'], stack.masters['c_rm'].tensor, rtol=1e-4, atol=1e-6)
        assert torch.allclose(got[r]['rv'], stack.masters['c_rv'].tensor, rtol=1e-4, atol=1e-6)
        ref = model.pc_decoder[1].flows[0].nvp2.T_mu_0_cond_w[0].weight.grad
        assert torch.allclose(got[r]['pgrad'], ref, rtol=1e-3, atol=1e-5)
    assert torch.equal(got[0]['dw0'], got[1]['dw0'])


_free_port_shared = _free_port()
