"""R ranks: time of one statistic exchange (gwtf_exchange_sum, the stand-alone kernel) issued back to back, and of a whole
synchronised forward + backward through the C ABI drivers against the same drivers on one rank's data alone.
torchrun --nproc-per-node R tools/exchange_probe.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from go_with_the_flows_b200 import _native as nat
from go_with_the_flows_b200 import flowstack

rank, lr = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev)
h = flowstack.peer_exchange(2048, dev)
assert h, 'no peer exchange'
lib = nat.lib()
st = torch.cuda.current_stream().cuda_stream
for n in (64, 592, 1184):
    data = torch.ones(n, dtype=torch.float64, device=dev)
    for _ in range(20):
        nat.check(lib.gwtf_exchange_sum(ctypes.c_void_p(h), nat.ptr(data), n, ctypes.c_void_p(st)), 'x')
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 500
    e0.record()
    for _ in range(reps):
        data.fill_(1.0)
        nat.check(lib.gwtf_exchange_sum(ctypes.c_void_p(h), nat.ptr(data), n, ctypes.c_void_p(st)), 'x')
    e1.record()
    torch.cuda.synchronize()
    t_x = e0.elapsed_time(e1) / reps
    e0.record()
    for _ in range(reps):
        data.fill_(1.0)
    e1.record()
    torch.cuda.synchronize()
    t_f = e0.elapsed_time(e1) / reps
    if rank == 0:
        print('n=%d doubles: %.2f us per exchange (fill + exchange %.2f us, fill alone %.2f us), total %.0f' %
              (n, (t_x - t_f) * 1e3, t_x * 1e3, t_f * 1e3, float(data[0])), flush=True)
dist.barrier()
dist.destroy_process_group()
