"""Per-kernel GPU time of one train step (torch profiler, CUDA activities only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from torch.profiler import profile, ProfilerActivity
from go_with_the_flows_b200.networks.losses import FlowMixtureNLL
cfg, model = bench.build_model('generative', 'cuda')
model.train(); model.mode = 'training'
p, g = bench.synthetic(64, 2048, 128); p, g = p.cuda(), g.cuda()
loss = FlowMixtureNLL()
params = list(model.parameters())
def step():
    for q in params: q.grad = None
    gg = g.detach().requires_grad_(True)
    out, logits = model.decode(p, gg, 2048)
    loss(out, logits).backward()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
ev = prof.key_averages()
rows = sorted(((e.device_time_total / 3e3, e.count // 3, e.key) for e in ev if e.device_time_total > 0 and e.device_type.name == 'CUDA'), reverse=True)
tot = sum(r[0] for r in rows)
print('GPU busy per step %.2f ms' % tot)
for ms, n, k in rows[:28]:
    print('%8.3f ms %5d  %s' % (ms, n, k[:110]))
