"""ctypes binding of libdesman_b200.so (include/desman_b200.h).

There is no CPU fallback: if the CUDA library is missing or no device is visible every
entry point raises.  Build the library with `python -m desman_b200.build` (or
`__graft_entry__.build()`).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DESMAN_B200_LIB: an instrumented build of the same library (tools/kprof.py); never a different implementation
LIB_PATH = os.environ.get("DESMAN_B200_LIB") or os.path.join(_HERE, "libdesman_b200.so")

RNG_MT19937, RNG_PHILOX = 0, 1
MAX_G = 32
K_NAMES = ("tau_sample", "mu_stats", "draw_gamma_eta", "finalize", "mt19937", "nmft", "other", "tau_group", "maintain", "tau_update")

_p64 = C.POINTER(C.c_int64)
_pd = C.POINTER(C.c_double)
_pu8 = C.POINTER(C.c_uint8)
_ctx = C.c_void_p

# every symbol include/desman_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "c_initRNG": (None, []),
    "c_setRNG": (None, [C.c_ulong]),
    "c_freeRNG": (None, []),
    "c_sample_tau": (C.c_int, [_p64, _pd, _pd, _p64, C.c_int, C.c_int, C.c_int]),
    "desman_last_error": (C.c_char_p, []),
    "desman_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "desman_build_info": (C.c_char_p, []),
    "desman_ctx_create": (C.c_int, [C.POINTER(_ctx), C.c_int, C.c_uint64, C.c_int]),
    "desman_ctx_destroy": (C.c_int, [_ctx]),
    "desman_set_counts": (C.c_int, [_ctx, _p64, C.c_int64, C.c_int, C.c_int64, C.c_int64]),
    "desman_set_hyper": (C.c_int, [_ctx, C.c_double, C.c_double, C.c_double]),
    "desman_set_rng": (C.c_int, [_ctx, C.c_uint64, C.c_uint32, C.c_uint64]),
    "desman_get_rng": (C.c_int, [_ctx, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "desman_set_state": (C.c_int, [_ctx, _p64, _pd, _pd, C.c_int]),
    "desman_get_state": (C.c_int, [_ctx, _p64, _pd, _pd]),
    "desman_set_tau_index": (C.c_int, [_ctx, _pu8, C.c_int]),
    "desman_get_tau_index": (C.c_int, [_ctx, _pu8]),
    "desman_sample_tau": (C.c_int, [_ctx, _p64]),
    "desman_mu_stats": (C.c_int, [_ctx, _p64, _p64]),
    "desman_draw_gamma_eta": (C.c_int, [_ctx, _p64, _p64, _pd, _pd]),
    "desman_loglik": (C.c_int, [_ctx, _pd, _pd]),
    "desman_loglik_general": (C.c_int, [_ctx, _pd, _pd, _pd, C.c_int, _pd]),
    "desman_state_logprob": (C.c_int, [_ctx, _p64, C.c_int64, C.c_int, _pd, _pd, C.c_int, _p64, _pd, _pd, _pd, _pd, _p64]),
    "desman_update": (C.c_int, [_ctx, C.c_int, _pd, _pd, _pd, _pd, _p64]),
    "desman_update_tau": (C.c_int, [_ctx, C.c_int, _pd, _pd, _pd, _pd, _p64]),
    "desman_get_star": (C.c_int, [_ctx, _p64, _pd, _pd, _pd, C.POINTER(C.c_int)]),
    "desman_get_star_index": (C.c_int, [_ctx, _pu8]),
    "desman_get_tau_sum": (C.c_int, [_ctx, _p64]),
    "desman_get_tau_sum_u32": (C.c_int, [_ctx, C.POINTER(C.c_uint32)]),
    "desman_nmft_factorize": (C.c_int, [_ctx, _p64, C.c_int64, C.c_int, C.c_int, _pd, _pd, C.c_int, C.c_double, C.c_int,
                                        C.POINTER(C.c_int), _pd, _pd]),
    "desman_set_option": (C.c_int, [_ctx, C.c_char_p, C.c_int64]),
    "desman_get_tier_counts": (C.c_int, [_ctx, _p64, C.c_int]),
    "desman_get_group_stats": (C.c_int, [_ctx, _p64]),
    "desman_get_esum_store": (C.c_int, [_ctx, _p64]),
    "desman_batch_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(_p64), C.POINTER(C.c_int), C.c_int]),
    "desman_batch_sample_tau": (C.c_int, [C.c_void_p, C.POINTER(_p64), C.POINTER(_pd), _pd, C.c_int, C.POINTER(C.c_int)]),
    "desman_batch_destroy": (C.c_int, [C.c_void_p]),
    "desman_sample_tau_fix": (C.c_int, [_ctx, C.c_int, _pd, C.POINTER(C.c_int64)]),
    "desman_debug_screen": (C.c_int, [_ctx, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]),
    "desman_nmft_last_timing": (C.c_int, [_pd, C.POINTER(C.c_int)]),
    "desman_comm_kind": (C.c_int, [_ctx]),
    "desman_comm_unique_id": (C.c_int, [C.c_char_p]),
    "desman_comm_init": (C.c_int, [_ctx, C.c_char_p, C.c_int, C.c_int]),
    "desman_set_profiling": (C.c_int, [_ctx, C.c_int, C.c_int]),
    "desman_get_timing": (C.c_int, [_ctx, _pd, _pd, _p64]),
    "desman_synchronize": (C.c_int, [_ctx]),
}

_LIB = None


class DesmanB200Error(RuntimeError):
    pass


def lib():
    """Load libdesman_b200.so; raises if it has not been built (no fallback path exists)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise DesmanB200Error(
                "libdesman_b200.so is not built (%s). Run `python -m desman_b200.build`; "
                "desman_b200 has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def last_error():
    return lib().desman_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != 0:
        raise DesmanB200Error("%s failed (%d): %s" % (what, rc, last_error()))


def device_count():
    n = C.c_int(0)
    check(lib().desman_device_count(C.byref(n)), "desman_device_count")
    return n.value


def as_i64(a, name="array"):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_p64)


def as_f64(a, name="array"):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_pd)


def ptr_d(a):
    return a.ctypes.data_as(_pd) if a is not None else None


def ptr_i64(a):
    return a.ctypes.data_as(_p64) if a is not None else None
