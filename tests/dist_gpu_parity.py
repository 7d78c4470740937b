"""torchrun --nproc-per-node 2 tests/dist_gpu_parity.py
Two ranks (one per GPU), 2 shapes each, train-mode fwd+bwd with SyncBN statistics exchanged between
kernel phases and flat gradient all-reduce: must equal the fp64 golden result of the reference on
the concatenated batch of 4."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from tests.util import Golden, build_dropin, max_rel, rel_l2


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl')
    from go_with_the_flows_b200.networks.losses import FlowMixtureNLL
    gd = Golden('small_freevar_global')
    per = 4 // world
    sl = slice(rank * per, (rank + 1) * per)
    model = build_dropin(gd, 'cuda')
    # as the reference does before DDP (train_ae.py:152): latent-side BatchNorms (p_prior, weights
    # encoder) become SyncBatchNorm; the decoder's own statistics are synchronised by our kernels
    model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model).train()
    model.mode = 'training'
    p = gd.t('in/p', torch.float32, 'cuda')[sl].contiguous().requires_grad_(True)
    g = gd.t('in/g', torch.float32, 'cuda')[sl].contiguous().requires_grad_(True)
    out, logits = model.decode(p, g, p.shape[2])
    nll = out[0]['mixture_nll']
    FlowMixtureNLL()(out, logits).backward()
    errs = {'nll': max_rel(nll.detach().cpu(), gd.t('train/nll')[sl]),
            'dp': rel_l2(p.grad.cpu() / world, gd.t('train/dp')[sl]),
            'dg': rel_l2(g.grad.cpu() / world, gd.t('train/dg')[sl])}
    named = dict(model.named_parameters())
    num = den = 0.0
    for k in gd.keys('train/grad/'):
        if not k.startswith('pc_decoder'):
            continue
        ref = gd.t(f'train/grad/{k}')
        got = named[k].grad.detach().cpu().double()
        num += float((got - ref).pow(2).sum())
        den += float(ref.pow(2).sum())
    errs['dparams'] = (num / den) ** 0.5
    sd = model.state_dict()
    errs['bn'] = max(max_rel(sd[k].cpu(), gd.t(f'train/bn/{k}'), floor=1e-3) for k in gd.keys('train/bn/')
                     if k.startswith('pc_decoder') and not k.endswith('num_batches_tracked'))
    ok = errs['nll'] < 1e-4 and errs['dp'] < 1e-4 and errs['dg'] < 1e-4 and errs['dparams'] < 1e-4 and errs['bn'] < 1e-4
    print('rank', rank, 'OK' if ok else 'FAIL', errs, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
