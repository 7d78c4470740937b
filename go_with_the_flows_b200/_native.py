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


ENGINE_DEFAULT, ENGINE_FMA, ENGINE_TC_FWD, ENGINE_MMA, ENGINE_TC = -1, 0, 2, 3, 4
ENGINE_NAMES = {ENGINE_FMA: 'fma', ENGINE_TC_FWD: 'tcgen05_fwd', ENGINE_MMA: 'mma', ENGINE_TC: 'tcgen05'}
FLAG_NO_PDL = 1
PRECISION_3XTF32, PRECISION_TF32 = 0, 1


class StackDesc(ctypes.Structure):
    """gwtf_stack_desc (include/gwtf.h): geometry + the caller-owned execution options."""
    _fields_ = [('n_components', c_i), ('n_layers', c_i), ('n_features', c_i), ('rec_stride', c_i),
                ('warp_mask', ctypes.c_uint8 * MAX_LAYERS),
                ('engine', c_i), ('flags', c_i), ('eval_precision', c_i), ('reserved', c_i),
                ('exchange', ctypes.c_void_p), ('nonfinite', ctypes.c_void_p)]


_D = ctypes.POINTER(StackDesc)
c_u64 = ctypes.c_uint64
c_i64 = ctypes.c_int64

# name -> argtypes; every function returns int (0 = ok) unless listed in _RESTYPES
_SIGNATURES = {
    'gwtf_version': [],
    'gwtf_resolved_engine': [_D, c_i],
    'gwtf_rec_stride': [c_i],
    'gwtf_param_offsets': [c_i, c_i, ctypes.POINTER(c_i), ctypes.POINTER(c_i)],
    'gwtf_fma_peak_tflops': [c_i, ctypes.POINTER(c_d), c_f],
    'gwtf_mma_peak_tflops': [c_i, ctypes.POINTER(c_d), c_f],
    'gwtf_nll_fwd_eval': [_D, c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_f, c_f, c_f, c_f, c_f],
    'gwtf_nll_fwd_eval_layers': [_D] + [c_f] * 7 + [c_i64, c_i, c_i, c_f, c_f, c_f],
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
    'gwtf_mixture_cdf': [c_f, c_i, c_i, c_f, c_f],
    'gwtf_sample_layers': [_D, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_u64, ctypes.c_uint32, c_f, c_f, c_f, c_i64,
                           c_f, c_f, c_f, c_f],
    'gwtf_exchange_create': [c_i, c_i, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), c_i, c_d,
                             ctypes.POINTER(ctypes.c_void_p)],
    'gwtf_exchange_destroy': [c_f],
    'gwtf_exchange_world': [c_f],
    'gwtf_exchange_resync': [c_f, c_u64],
    'gwtf_exchange_sum': [c_f, c_f, c_i, c_f],
    'gwtf_fwd_all_ranks': [_D, c_i] + [c_f] * 13 + [c_i, c_i, c_f, c_f, c_d, c_f],
    'gwtf_bwd_all_ranks': [_D, c_i] + [c_f] * 22 + [c_i, c_i, c_d, c_f],
    'gwtf_sample': [_D, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_u64, ctypes.c_uint32, c_f, c_f,
                    c_f, c_f, c_f, c_f],
    'gwtf_adam_step': [c_f, c_f, c_f, c_f, c_f, c_i64, c_d, c_d, c_d, c_d, c_d, c_i64, c_f],
    'gwtf_debug_tile_schedule': [c_i, c_i, c_i, c_i, c_f, c_f, c_f],
}
_RESTYPES = {
    'gwtf_last_error_string': ([], ctypes.c_char_p),
    'gwtf_keep_floats': ([_D, c_i, c_i], c_i64),
    'gwtf_eval_layers_workspace_bytes': ([_D, c_i, c_i], c_i64),
    'gwtf_sample_workspace_bytes': ([_D, c_i, c_i], c_i64),
    'gwtf_exchange_seq': ([c_f], c_u64),
}

EXPORTED = sorted(_SIGNATURES) + sorted(_RESTYPES)

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
        for name, (argtypes, restype) in _RESTYPES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().gwtf_last_error_string().decode('utf-8', 'replace')
        raise GwtfError('%s failed (rc=%d): %s' % (what, rc, msg))


def ptr(t):
    """Device pointer of a tensor (or NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())
