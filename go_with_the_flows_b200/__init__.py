"""go_with_the_flows_b200 -- B200-native mixture-of-normalizing-flows hot path.

Drop-in for the flow modules of janisgp/go_with_the_flows (`lib/networks`): same class names,
constructor / forward signatures and state_dict keys, with the per-point mixture-of-flows
log-likelihood (forward + backward) and the sampling pass running in hand-written sm_100a
CUDA kernels behind the C ABI of `libgwtf.so` (see include/gwtf.h).  There is no CPU fallback:
using a flow module without the built extension raises.
"""
from . import networks  # noqa: F401

__all__ = ['networks']
