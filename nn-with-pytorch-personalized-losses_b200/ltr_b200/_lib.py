"""ctypes binding of the C-ABI in include/ltr_b200.h.

There is NO fallback: if the CUDA library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_uint, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LTR_B200_LIB") or os.path.join(_HERE, "lib", "libltr_b200.so")   # (override: A/B builds)

WORKSPACE_BYTES = 262144
RISK_XCHG_BYTES = 16384
RISK_MAX_WORLD = 16

_P = c_void_p
_SIGNATURES = {
    # name: (restype, [argtypes])
    "ltr_version": (c_char_p, []),
    "ltr_last_error": (c_char_p, []),
    "ltr_max_slate_len": (c_int, []),
    "ltr_fill_disc_prefix": (c_int, [_P, c_int, _P]),
    "ltr_approx_ndcg_scratch_bytes": (c_size_t, [c_int, c_int]),
    "ltr_approx_ndcg_fwdbwd": (c_int, [_P, _P, c_int, c_int, c_float, c_float, c_float, _P, c_int, _P, _P, _P, _P,
                                       c_size_t, _P, _P]),
    "ltr_approx_ndcg_host_scratch_bytes": (c_size_t, [c_int, c_int]),
    "ltr_approx_ndcg_fwdbwd_host": (c_int, [_P, _P, c_int, c_int, c_float, c_float, c_float, _P, c_int, _P, _P, _P, _P,
                                            c_size_t, _P, _P]),
    "ltr_listnet_fwdbwd": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P]),
    "ltr_ordinal_fwdbwd": (c_int, [_P, _P, c_int, c_int, c_int, c_float, _P, _P, _P, _P, _P]),
    "ltr_sort_desc": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P]),
    "ltr_ndcg_at_k": (c_int, [_P, _P, c_int, c_int, POINTER(c_int), c_int, c_int, _P, _P, _P, _P]),
    "ltr_ndcg_at_k_scratch_bytes": (c_size_t, [c_int, c_int]),
    "ltr_ndcg_at_k_ws": (c_int, [_P, _P, c_int, c_int, POINTER(c_int), c_int, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "ltr_ndcg_at_k_host_scratch_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ltr_ndcg_at_k_host": (c_int, [_P, _P, c_int, c_int, POINTER(c_int), c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "ltr_ndcg_at_k_f64": (c_int, [_P, _P, c_int, c_int, POINTER(c_int), c_int, c_int, c_int, c_double, _P, _P, _P, _P]),
    "ltr_dcg_at_k_f64": (c_int, [_P, _P, c_int, c_int, POINTER(c_int), c_int, c_int, c_int, _P, _P, _P, _P]),
    "ltr_exact_ndcg": (c_int, [_P, _P, c_int, c_int, _P, _P, _P]),
    "ltr_order_score_fwdbwd": (c_int, [_P, _P, c_int, c_int, _P, _P, _P, _P]),
    "ltr_lambda_fwdbwd": (c_int, [_P, _P, c_int, c_int, c_float, c_float, c_int, c_int, c_float, c_float, c_int, c_int,
                                  _P, _P, _P, _P, _P, _P, _P]),
    "ltr_risk_lambda_rows": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "ltr_risk_lambda_rows_bwd": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, _P, _P]),
    "ltr_lambda_pair_matrix": (c_int, [_P, _P, c_int, c_int, c_float, c_float, c_int, c_int, c_float, c_float, c_int, _P,
                                       _P, _P]),
    "ltr_lambda_pair_matrix_bwd": (c_int, [_P, _P, c_int, c_int, c_float, c_float, c_int, c_int, c_float, c_float, c_int,
                                           _P, _P, _P]),
    "ltr_risk_phase": (c_int, [c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, _P, c_int, c_int,
                               _P, _P, _P, _P]),
    "ltr_risk_listnet_fwdbwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                        c_float, _P, _P, _P, _P, c_int, _P, c_int, c_int, POINTER(c_void_p), _P, _P]),
    "ltr_risk_xchg_alloc": (c_int, [POINTER(c_void_p)]),
    "ltr_risk_xchg_free": (c_int, [_P]),
    "ltr_ipc_export": (c_int, [_P, _P]),
    "ltr_ipc_open": (c_int, [_P, POINTER(c_void_p)]),
    "ltr_ipc_close": (c_int, [_P]),
    "ltr_risk_xchg_status": (c_int, [_P, POINTER(c_int)]),
    "ltr_georisk_default_f64": (c_int, [_P, c_int, c_int, c_double, _P, _P]),
    "ltr_triple_scratch_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ltr_triple_approx_ndcg_fwdbwd": (c_int, [_P, _P, _P, _P, _P, c_float, c_int, c_int, c_int, c_float, c_float, c_float, _P,
                                              c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P, _P]),
    "ltr_triple_step_scratch_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ltr_triple_approx_ndcg_step": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, c_float, c_float, _P, c_int,
                                            _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P, _P]),
    "ltr_triple_listnet_step": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_int,
                                        _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P, _P]),
    "ltr_peer_xchg_bytes": (c_size_t, []),
    "ltr_peer_xchg_max_floats": (c_int, []),
    "ltr_peer_xchg_alloc": (c_int, [POINTER(_P)]),
    "ltr_peer_xchg_free": (c_int, [_P]),
    "ltr_peer_xchg_status": (c_int, [_P, POINTER(c_int)]),
    "ltr_peer_allreduce_avg": (c_int, [POINTER(_P), POINTER(c_int), c_int, _P, POINTER(_P), c_int, c_int, _P]),
    "ltr_svmlight_scan": (c_int, [c_char_p, c_int, POINTER(c_longlong)]),
    "ltr_svmlight_load": (c_int, [c_char_p, c_longlong, c_longlong, c_int, c_int, c_int, c_float, _P, _P, _P, _P, c_int]),
    "ltr_baseline_csv_scan": (c_int, [c_char_p, POINTER(c_longlong)]),
    "ltr_baseline_csv_load": (c_int, [c_char_p, c_longlong, c_longlong, c_int, c_int, _P, _P]),
    "ltr_probe_sfu": (c_int, [_P, c_int, POINTER(c_longlong), _P]),
    "ltr_probe_fma": (c_int, [_P, c_int, POINTER(c_longlong), _P]),
    "ltr_probe_tma_stream": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P]),
}

_lib = None


class LtrError(RuntimeError):
    pass


def load():
    """Load libltr_b200.so (once).  Raises if it has not been built -- there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LtrError(
            f"{LIB_PATH} not found: build it with `python -m ltr_b200.build` "
            "(or __graft_entry__.build()). The B200 path has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return list(_SIGNATURES.keys())


def check(status: int, what: str):
    if status != 0:
        msg = load().ltr_last_error().decode(errors="replace")
        raise LtrError(f"{what} failed with status {status}: {msg}")
