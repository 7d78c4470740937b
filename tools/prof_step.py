"""One C2 train-mode step (64 x 2048, airplane model) for profiling: `python tools/prof_step.py [n_steps]`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from go_with_the_flows_b200.networks.losses import FlowMixtureNLL

n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg, model = bench.build_model('generative', 'cuda')
model.train()
model.mode = 'training'
p, g = bench.synthetic(64, 2048, 128)
p, g = p.cuda(), g.cuda()
loss = FlowMixtureNLL()
params = list(model.parameters())


def step():
    for q in params:
        q.grad = None
    gg = g.detach().requires_grad_(True)
    out, logits = model.decode(p, gg, 2048)
    loss(out, logits).backward()


for _ in range(n_steps):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
step()
e1.record()
torch.cuda.synchronize()
print('step ms', e0.elapsed_time(e1))
