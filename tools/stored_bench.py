import sys, os
sys.path.insert(0, '/root/repo')
import torch, bench
from go_with_the_flows_b200.networks.losses import FlowMixtureNLL
cfg, model = bench.build_model('generative', 'cuda')
model.train(); model.mode = 'training'
model.flow_stack().keep_activations = True
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): step()
e1.record(); torch.cuda.synchronize()
print('stored-y1 step ms', e0.elapsed_time(e1) / 3)
