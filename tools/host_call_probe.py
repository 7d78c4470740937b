"""Host time of the single-call drivers (is anything in them synchronous?)."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from go_with_the_flows_b200 import _native as nat
from go_with_the_flows_b200.networks.losses import FlowMixtureNLL
lib = nat.lib()
cfg, model = bench.build_model('generative', 'cuda')
model.train(); model.mode = 'training'
p, g = bench.synthetic(64, 2048, 128); p, g = p.cuda(), g.cuda()
loss = FlowMixtureNLL()
params = list(model.parameters())
import go_with_the_flows_b200.flowstack as fs
T = {}
def wrap(name):
    fn = getattr(lib, name)
    def w(*a):
        t = time.perf_counter(); r = fn(*a); T.setdefault(name, []).append((time.perf_counter() - t) * 1e3); return r
    return w
class L:  # proxy
    def __getattr__(self, n):
        return wrap(n) if n in ('gwtf_fwd_all', 'gwtf_bwd_all') else getattr(lib, n)
proxy = L()
fs.nat.lib = lambda: proxy
def step():
    for q in params: q.grad = None
    gg = g.detach().requires_grad_(True)
    t0 = time.perf_counter()
    out, logits = model.decode(p, gg, 2048)
    l = loss(out, logits)
    t1 = time.perf_counter()
    l.backward()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3
for _ in range(3): step()
T.clear()
rows = [step() for _ in range(5)]
print('host fwd %.2f ms, host bwd %.2f ms, drain %.2f ms' % tuple(sum(r[i] for r in rows) / 5 for i in range(3)))
for k, v in T.items(): print(k, 'host ms per call %.2f' % (sum(v) / len(v)))
