"""Kernel-only and whole-step times of the C2 workload for each engine x keep-activations setting."""
import sys, os
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
def step():
    for q in params: q.grad = None
    gg = g.detach().requires_grad_(True)
    out, logits = model.decode(p, gg, 2048)
    loss(out, logits).backward()
combos = [tuple(int(c) for c in a.split(',')) for a in sys.argv[1:]] or [(3, 1), (3, 0), (2, 0)]
for eng, keep in combos:
    lib.gwtf_set_tensor_cores(eng)
    model.flow_stack().keep_activations = bool(keep)
    for _ in range(3): step()
    k = bench.kernel_only_times(model, p, g, 3, keep=bool(keep))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): step()
    e1.record(); torch.cuda.synchronize()
    print('engine %d keep %d: step %.2f ms | fwd %.2f bwd %.2f | per launch us: %s' % (
        eng, keep, e0.elapsed_time(e1) / 5, k['fwd_ms'], k['bwd_ms'],
        {n[:-14]: round(1e3 * v, 1) for n, v in k.items() if n.endswith('per_launch')}), flush=True)
