"""torchrun --nproc-per-node R tests/dist_gpu_ddp.py
The wrapping the reference's training CLI applies (train_ae.py:151-153):

    model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    model = DistributedDataParallel(model, device_ids=[rank], find_unused_parameters=True)

then N optimizer steps with the custom Adam (lib/networks/optimizers.py) through the calls training.py:40-54
makes -- `model(g_clouds, p_clouds, ...)`, `Flow_Mixture_Loss`, `loss.backward()`, `optimizer.step()`.
R ranks on the sharded batch must reproduce ONE process on the concatenated batch: losses per step and every
parameter after the last step (the single-process path is itself pinned to the reference by the golden tests).
"""
import ast
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from tests.util import GOLDEN_DIR

STEPS = 3


def build(z, cfg, dev):
    from go_with_the_flows_b200.networks.flow_mixture import Flow_Mixture_Model
    model = Flow_Mixture_Model(**cfg)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd/')}
    model.load_state_dict({k: (v.float() if v.is_floating_point() else v) for k, v in sd.items()}, strict=True)
    model = model.to(dev).train()
    model.mode = 'training'
    return model


def run(model, core, g_in, p_in, eps, world):
    """`model` is what gets called (DDP wrapper or the bare module), `core` the module that owns reparameterize."""
    from go_with_the_flows_b200.networks.losses import Flow_Mixture_Loss
    from go_with_the_flows_b200.networks.optimizers import Adam
    cfg = core._test_cfg
    loss_fn = Flow_Mixture_Loss(**cfg)
    opt = Adam(model.parameters(), lr=2e-3, betas=(0.9, 0.995), weight_decay=1e-5, amsgrad=True)
    losses = []
    for step in range(STEPS):
        core.reparameterize = lambda mu, logvar, _e=eps[step]: _e * torch.exp(0.5 * logvar) + mu
        out_enc, out_dec, logits = model(g_in, p_in, images=None, n_sampled_points=None, labeled_samples=False,
                                         warmup=False)
        loss, pnll, gnll, gent = loss_fn(out_enc, out_dec, logits)
        opt.zero_grad()
        loss.backward()
        grads = {k: (v.grad.detach().clone() if v.grad is not None else torch.zeros_like(v))
                 for k, v in core.named_parameters()}
        opt.step()
        t = torch.stack([loss.detach(), pnll.detach(), gnll.detach(), gent.detach()])
        if world > 1:
            dist.all_reduce(t)
            t /= world
        losses.append(t.cpu())
    return torch.stack(losses), {k: v.detach().clone() for k, v in core.state_dict().items()}, grads


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=dev)
    z = np.load(os.path.join(GOLDEN_DIR, 'model_small.npz'))
    cfg = ast.literal_eval(str(z['meta']))
    per = 2
    Bt, N, G = per * world, 50, cfg['g_latent_space_size']
    gen = torch.Generator().manual_seed(99)
    g_all = (0.4 * torch.randn(Bt, 3, N, generator=gen)).to(dev)
    p_all = (0.4 * torch.randn(Bt, 3, N, generator=gen)).to(dev)
    eps_all = torch.randn(STEPS, Bt, G, generator=gen).to(dev)
    sl = slice(rank * per, (rank + 1) * per)

    # ---- R ranks, wrapped exactly like train_ae.py:151-153
    model = build(z, cfg, dev)
    model._test_cfg = cfg
    model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], find_unused_parameters=True)
    losses_r, sd_r, grads_r = run(ddp, ddp.module, g_all[sl].contiguous(), p_all[sl].contiguous(),
                                  eps_all[:, sl].contiguous(), world)
    # DDP invariant: every rank holds the same parameters after the steps
    chk = torch.stack([v.double().sum() for v in sd_r.values() if v.is_floating_point()]).sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(float((hi - lo).abs()) <= 1e-9 * max(1.0, float(hi.abs())))

    # ---- one process on the concatenated batch (rank 0 only; no process group involved)
    ok = True
    if rank == 0:
        from go_with_the_flows_b200 import flowstack
        ref = build(z, cfg, dev)
        ref._test_cfg = cfg
        real_world = flowstack._world
        flowstack._world = lambda: 1                      # the single-process path
        try:
            losses_1, sd_1, grads_1 = run(ref, ref, g_all, p_all, eps_all, 1)
        finally:
            flowstack._world = real_world
        # gradients of the last step, norm-wise per parameter group (three Adam steps in, so earlier steps count too)
        groups = {'decoder': 'pc_decoder', 'encoder': ('pc_encoder', 'g_posterior'), 'prior': ('g_prior', 'g0_prior'),
                  'heads': ('p_prior', 'mixture_weights')}
        gerr = {}
        for name, pref in groups.items():
            num = sum(float((grads_r[k] - grads_1[k]).double().pow(2).sum()) for k in grads_1 if k.startswith(pref))
            den = sum(float(grads_1[k].double().pow(2).sum()) for k in grads_1 if k.startswith(pref))
            gerr[name] = (num / max(den, 1e-300)) ** 0.5
        lerr = float(((losses_r - losses_1).abs() / losses_1.abs().clamp_min(1.0)).max())
        ok = same and lerr < 1e-4 and all(v < 2e-3 for v in gerr.values())
        print('ddp-vs-single world=%d: ranks hold identical parameters: %s, loss err %.2e, last-step gradient err %s -> %s' %
              (world, same, lerr, {k: float('%.2e' % v) for k, v in gerr.items()}, 'OK' if ok else 'FAIL'), flush=True)
        print('losses (R ranks):', losses_r[:, 0].tolist(), ' single:', losses_1[:, 0].tolist(), flush=True)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if float(flag) > 0 else 1)


if __name__ == '__main__':
    main()
