"""One train-mode fwd+bwd step of the airplane decoder at B clouds x N points (python tools/step_sizes.py B [N]); prints the
loss and a gradient checksum.  Used to sweep batch sizes whose tile ranges start at odd offsets inside a shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from go_with_the_flows_b200.networks.losses import FlowMixtureNLL

B = int(sys.argv[1])
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
cfg, model = bench.build_model('generative', 'cuda')
model.train()
model.mode = 'training'
p, g = bench.synthetic(B, N, 128)
p, g = p.cuda(), g.cuda()
loss = FlowMixtureNLL()
gg = g.detach().requires_grad_(True)
out, logits = model.decode(p, gg, N)
l = loss(out, logits)
l.backward()
torch.cuda.synchronize()
print('B=%d N=%d loss %.6f dg %.6e' % (B, N, float(l), float(gg.grad.double().abs().sum())), flush=True)
