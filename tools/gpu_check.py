"""Print every parity metric of the CUDA path against the golden vectors (run under gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import GOLDEN_CASES, Golden
from tests import parity

torch.manual_seed(0)
for case in GOLDEN_CASES:
    gd = Golden(case)
    print('==', case, gd.meta)
    try:
        print('  eval fused nll err', parity.dropin_eval_fused_error(gd))
    except Exception as e:
        print('  eval fused FAILED', repr(e))
    for tag in ('eval', 'train'):
        print('  oracle fp32 noise', tag, parity.oracle_fp32_errors(gd, tag))
        for fused in (True, False):
            try:
                print('  ours', tag, 'fused' if fused else 'lists', parity.dropin_nll_errors(gd, tag, fused))
            except Exception as e:
                import traceback; traceback.print_exc()
                print('  ours', tag, fused, 'FAILED', repr(e))
    try:
        print('  sample', parity.dropin_sample_errors(gd))
    except Exception as e:
        import traceback; traceback.print_exc()
        print('  sample FAILED', repr(e))
torch.cuda.synchronize()
print('done')
