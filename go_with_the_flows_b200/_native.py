"""ctypes binding of libgwtf.so (the C ABI declared in include/gwtf.h)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libgwtf.so')

MAX_LAYERS = 96
MOM_STRIDE = 16

c_f = ctypes.c_void_p      # device pointers travel as integers
c_i = ctypes.c_int32
c_d = ctypes.c_double


class StackDesc(ctypes.Structure):
    _fields_ = [('n_components', c_i), ('n_layers', c_i), ('n_features', c_i), ('rec_stride', c_i),
                ('warp_mask', ctypes.c_uint8 * MAX_LAYERS)]


_D = ctypes.POINTER(StackDesc)

# name -> argtypes; every function returns int (0 = ok)
_SIGNATURES = {
    'gwtf_version': [],
    'gwtf_set_tensor_cores': [c_i],
    'gwtf_engine': [],
    'gwtf_set_pdl': [c_i],
    'gwtf_rec_stride': [c_i],
    'gwtf_param_offsets': [c_i, c_i, ctypes.POINTER(c_i), ctypes.POINTER(c_i)],
    'gwtf_fma_peak_tflops': [c_i, ctypes.POINTER(c_d), c_f],
    'gwtf_mma_peak_tflops': [c_i, ctypes.POINTER(c_d), c_f],
    'gwtf_nll_fwd_eval': [_D, c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_f, c_f, c_f, c_f, c_f],
    'gwtf_nll_fwd_eval_layers': [_D] + [c_f] * 8 + [c_i, c_i, c_f, c_f, c_f],
    'gwtf_fwd_moments': [_D, c_f, c_i, c_i, c_f, c_f],
    'gwtf_fwd_layer': [_D, c_i, c_i, c_i] + [c_f] * 10 + [c_i, c_i, c_d, c_f],
    'gwtf_fwd_layer_ex': [_D, c_i, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_i] + [c_f] * 8 + [c_i, c_i, c_d, c_f],
    'gwtf_fwd_bstat': [_D, c_f, c_f, c_f, c_d, c_f, c_f],
    'gwtf_nll_from_state': [_D, c_f, c_f, c_f, c_f, c_i, c_i, c_f, c_f, c_f],
    'gwtf_fwd_all': [_D, c_i] + [c_f] * 13 + [c_i, c_i, c_f, c_f, c_f],
    'gwtf_bwd_seed': [_D, c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_f, c_f, c_f, c_f, c_f],
    'gwtf_bwd_layer': [_D, c_i, c_i, c_i] + [c_f] * 14 + [c_i, c_i, c_d, c_f],
    'gwtf_bwd_finish': [_D, c_i, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_d, c_f],
    'gwtf_bwd_all': [_D, c_i] + [c_f] * 22 + [c_i, c_i, c_f],
    'gwtf_sample_plan': [_D, c_f, c_i, c_i, ctypes.c_uint64, ctypes.c_uint32, c_f, c_f, c_f, c_f],
    'gwtf_sample_layers': [_D, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, ctypes.c_uint64, ctypes.c_uint32, c_f, c_f, c_f,
                           c_f, c_f, c_f, c_f, c_f],
    'gwtf_exchange_attach': [c_i, c_i, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), c_i],
    'gwtf_exchange_world': [],
    'gwtf_exchange_sum': [c_f, c_i, c_f],
    'gwtf_fwd_all_ranks': [_D, c_i] + [c_f] * 13 + [c_i, c_i, c_f, c_f, c_d, c_f],
    'gwtf_bwd_all_ranks': [_D, c_i] + [c_f] * 22 + [c_i, c_i, c_d, c_f],
    'gwtf_sample': [_D, c_f, c_f, c_f, c_f, c_f, c_i, c_i, ctypes.c_uint64, ctypes.c_uint32, c_f, c_f,
                    c_f, c_f, c_f, c_f],
}

EXPORTED = sorted(_SIGNATURES) + ['gwtf_last_error_string', 'gwtf_keep_floats']

_lib = None


class GwtfError(RuntimeError):
    pass


def lib():
    """Load libgwtf.so or fail loudly -- the product path has no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GwtfError(
                'libgwtf.so is not built (%s). Run `python -c "import __graft_entry__ as g; g.build()"` '
                'or `python -m go_with_the_flows_b200.build`.' % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = ctypes.c_int
        handle.gwtf_last_error_string.argtypes = []
        handle.gwtf_last_error_string.restype = ctypes.c_char_p
        handle.gwtf_keep_floats.argtypes = [_D, c_i, c_i]
        handle.gwtf_keep_floats.restype = ctypes.c_int64
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().gwtf_last_error_string().decode('utf-8', 'replace')
        raise GwtfError('%s failed (rc=%d): %s' % (what, rc, msg))


def ptr(t):
    """Device pointer of a tensor (or NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())
