"""Cycles one compute thread of CTA (0, 0) spends in each stage of the tcgen05 backward kernels, summed over the 33
layers of one C2 step.  Needs a profiling build:
    GWTF_NVCC_EXTRA=-DGWTF_STAGE_CLOCKS python -m go_with_the_flows_b200.build --force
    python tools/stage_clocks.py            (on the GPU box)
    python -m go_with_the_flows_b200.build --force      (back to the product build)"""
import ctypes
import os
import runpy
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from go_with_the_flows_b200 import _native

lib = _native.lib()
sys.argv = [sys.argv[0], '2']
ns = runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'prof_step.py'))
buf = (ctypes.c_ulonglong * 32)()
lib.gwtf_debug_stage_clocks(None, 1)
ns['step']()
lib.gwtf_debug_stage_clocks(buf, 0)
names = {0: 'prologue (x, dO loads, operand row)', 1: 'wait A (y0, P)', 2: 'C0 relu/mask -> a0 operand', 3: 'wait B (y1)',
         4: 'C1a r -> D|P', 5: 'wait operand-buffer turn', 6: 'C1b copy TMEM -> smem', 7: 'wait C (da0)',
         8: 'C2 dy0 -> D|P + tile rows', 9: 'column sums', 10: 'wait D (du)', 11: 'du store', 12: 'per-shape restage',
         13: 'drain', 16: 'p0 prologue', 17: 'p0 wait A', 18: 'p0 C0', 19: 'p0 wait B', 20: 'p0 a1 -> operand',
         21: 'p0 wait C', 22: 'p0 head', 23: 'p0 sums', 24: 'p0 restage/flush', 25: 'p0 drain'}
for lo, hi, title in ((0, 16, 'phase 1'), (16, 32, 'phase 0')):
    tot = sum(buf[i] for i in range(lo, hi))
    if not tot:
        continue
    print('%s: %.1f us per launch at 1.9 GHz (33 launches)' % (title, tot / 33 / 1900.0))
    for i in range(lo, hi):
        if buf[i]:
            print('  %-40s %6.1f %%  %9.0f cycles/launch' % (names.get(i, str(i)), 100.0 * buf[i] / tot, buf[i] / 33))
