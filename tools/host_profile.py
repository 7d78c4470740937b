"""Where does the host time of one train step go?  (run under gpurun)"""
import cProfile, pstats, sys, os, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from go_with_the_flows_b200.networks.losses import FlowMixtureNLL

cfg, model = bench.build_model('generative', 'cuda')
model.train(); model.mode = 'training'
p, g = bench.synthetic(64, 2048, 128)
p, g = p.cuda(), g.cuda()
loss = FlowMixtureNLL()
params = list(model.parameters())

def step():
    for q in params: q.grad = None
    gg = g.detach().requires_grad_(True)
    out, logits = model.decode(p, gg, 2048)
    l = loss(out, logits)
    l.backward()
    return l

for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print('host issue time per step %.2f ms, incl. drain %.2f ms' % ((t1 - t0) / 5 * 1e3, (t2 - t0) / 5 * 1e3))
stack = model.flow_stack()
def t(fn, n=5):
    torch.cuda.synchronize(); a = time.perf_counter()
    for _ in range(n): r = fn()
    b = time.perf_counter(); torch.cuda.synchronize(); c = time.perf_counter()
    return (b - a) / n * 1e3, (c - a) / n * 1e3
print('pack_params  host/total ms', t(stack.pack_params))
print('pack_bn      host/total ms', t(stack.pack_bn))
print('film(train)  host/total ms', t(lambda: stack.film(g, True, False)))
with torch.no_grad():
    print('film(nograd) host/total ms', t(lambda: stack.film(g, False, False)))
pr = cProfile.Profile()
pr.enable()
for _ in range(3): step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(35)
print(s.getvalue()[:6000])
