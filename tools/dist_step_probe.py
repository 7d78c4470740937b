"""2+ ranks: where does a synchronised train step spend its time?  torchrun --nproc-per-node R tools/dist_step_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from go_with_the_flows_b200.networks.losses import FlowMixtureNLL
rank = int(os.environ['RANK']); lr = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl')
dev = torch.device('cuda', lr)
cfg, model = bench.build_model('generative', dev)
model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model) if False else model
model.train(); model.mode = 'training'
p, g = bench.synthetic(64, 2048, 128, seed_shift=rank); p, g = p.to(dev), g.to(dev)
loss = FlowMixtureNLL()
params = list(model.parameters())
marks = {}
import go_with_the_flows_b200.flowstack as fs
orig_fwd = fs._StackNLLPass.forward
def step(sync_each):
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for q in params: q.grad = None
    gg = g.detach().requires_grad_(True)
    out, logits = model.decode(p, gg, 2048)
    t1 = time.perf_counter()
    l = loss(out, logits)
    l.backward()
    t2 = time.perf_counter()
    e1.record()
    if sync_each:
        torch.cuda.synchronize()
    t3 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t0) * 1e3, e0, e1
for _ in range(3): step(True)
dist.barrier(); torch.cuda.synchronize()
for mode in (True, False):
    rows = []
    t_all = time.perf_counter()
    for _ in range(6): rows.append(step(mode))
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t_all) * 1e3 / 6
    dist.barrier()
    print('rank %d sync_each=%s: wall/step %.2f ms | host fwd %.2f bwd %.2f total %.2f | gpu %.2f' % (
        rank, mode, t_all, sum(r[0] for r in rows) / 6, sum(r[1] for r in rows) / 6, sum(r[2] for r in rows) / 6,
        sum(r[3].elapsed_time(r[4]) for r in rows) / 6), flush=True)
dist.destroy_process_group()
