"""Eval-mode NLL: fused single-launch FMA kernel vs the per-layer tensor-core kernels (train=0)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from go_with_the_flows_b200 import _native as nat
from go_with_the_flows_b200.flowstack import _stream_ptr
lib = nat.lib()
cfg, model = bench.build_model('generative', 'cuda')
model.eval()
stack = model.flow_stack(); stack.prepare()
K, L, Fd = stack.K, stack.L, stack.F
for B, N in ((64, 2048), (4, 2048), (256, 2048)):
    p, g = bench.synthetic(B, N, 128); p, g = p.cuda(), g.cuda()
    with torch.no_grad():
        film = stack.film(g, False, False); params = stack.pack_params().contiguous(); bnbuf = stack.pack_bn()
        mu_b, lv_b = model.base_gaussian(g); base = torch.stack([mu_b, lv_b], 1).contiguous()
        logits = model.get_weights(g); logw = (logits - torch.logsumexp(logits, -1, keepdim=True)).contiguous()
    nll1 = torch.empty(B, N, device='cuda'); nll2 = torch.empty(B, N, device='cuda')
    ubuf = torch.empty(L, K, B, 3, N, device='cuda'); ld = torch.zeros(K, B, N, device='cuda')
    desc = ctypes.byref(stack.desc); P = nat.ptr
    def fused():
        nat.check(lib.gwtf_nll_fwd_eval(desc, P(params), P(bnbuf), P(film), P(p), P(base), P(logw), B, N, P(nll1), None, None, None, _stream_ptr()), 'e')
    def layered():
        nat.check(lib.gwtf_fwd_all(desc, 0, P(params), P(bnbuf), P(film), P(p), P(base), P(logw), P(ubuf), P(ld), None, None, None, None, None, B, N, P(nll2), None, _stream_ptr()), 'f')
    def timed(fn):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 5
    for eng in (2, 3):
        lib.gwtf_set_tensor_cores(eng)
        tl = timed(layered)
        print('B=%d N=%d engine %d: fused %.3f ms, layered %.3f ms, max |dnll| %.2e' % (B, N, eng, timed(fused), tl, float((nll1 - nll2).abs().max())))
