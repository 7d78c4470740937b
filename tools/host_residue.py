"""Which kernels make up the step besides the 132 layer launches?  torch.profiler over a few C2 steps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from go_with_the_flows_b200.networks.losses import FlowMixtureNLL

cfg, model = bench.build_model('generative', 'cuda')
model.train()
model.mode = 'training'
p, g = bench.synthetic(64, 2048, 128)
p, g = p.cuda(), g.cuda()
step = bench.make_step(model, 1, 2048)
for _ in range(3):
    step(p, g)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step(p, g)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    dt = getattr(e, 'device_time_total', None)
    if dt is None:
        dt = getattr(e, 'cuda_time_total', 0)
    if dt > 0:
        rows.append((dt / 3.0, e.count / 3.0, e.key[:90]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print('device time per step (us), launches per step, name   [sum %.0f us]' % tot)
for r in rows[:45]:
    print('%9.1f %7.1f  %s' % r)
import time
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    step(p, g)
t_cpu = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print('host time to ENQUEUE a step: %.2f ms; wall per step %.2f ms' % (t_cpu / 10 * 1e3, t_all / 10 * 1e3))
