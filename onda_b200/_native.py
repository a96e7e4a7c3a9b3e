"""ctypes binding of ``libonda_b200.so`` (include/onda_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, an
exception is raised.  Nothing in this package computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libonda_b200.so")

OK = 0
METRIC = {"euclidean": 0, "mahalanobis": 1}
IMPL = {"auto": 0, "simt": 1, "tcgen05": 2}
NUM_STATS = 8
STAT_PROTO_CONF, STAT_PRIOR_CONF, STAT_PL_CONF, STAT_PL_PIXELS, STAT_PIXELS, STAT_ENTROPY = range(6)
IGNORE_LABEL = 255
REGULARIZER = {None: 0, "": 0, "MRKLD": 1, "MRENT": 2}
MAX_CLASSES = 32

_lib = None

_p = C.c_void_p
_SIGNATURES = {
    "onda_abi_version": (C.c_int, []),
    "onda_last_error": (C.c_char_p, []),
    "onda_sm_count": (C.c_int, []),
    "onda_launch_count": (C.c_ulonglong, []),
    "onda_set_tile_schedule": (C.c_int, [C.c_int]),
    "onda_debug_set_buffer": (C.c_int, [_p]),
    "onda_debug_load_probe": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p]),
    "onda_kernel_timing_enable": (C.c_int, [C.c_int]),
    "onda_kernel_timing_read": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "onda_table_floats": (C.c_size_t, [C.c_int, C.c_int]),
    "onda_sums_floats": (C.c_size_t, [C.c_int, C.c_int]),
    "onda_impl_supported": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "onda_fused_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "onda_build_distance_table": (C.c_int, [_p, _p, _p, C.c_int, C.c_int, C.c_int, _p, _p]),
    "onda_table_global_std": (C.c_int, [_p, C.c_int, C.c_int, _p, _p]),
    "onda_prototype_std": (C.c_int, [_p, _p, C.c_int, C.c_int, _p, _p]),
    "onda_pseudolabel_fused": (C.c_int, [_p, _p, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                         _p, _p, _p, _p, _p, C.c_size_t, C.c_int, _p]),
    "onda_pseudolabel_fused_guarded": (C.c_int, [_p, _p, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                                 _p, _p, _p, _p, _p, C.c_size_t, C.c_int, _p, _p, C.c_int, _p]),
    "onda_class_sums_labelled": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p, C.c_size_t, _p]),
    "onda_ema_update": (C.c_int, [_p, _p, _p, C.c_int, C.c_int, C.c_float, _p]),
    "onda_ema_update_and_table": (C.c_int, [_p, _p, _p, _p, C.c_int, C.c_int, C.c_float, C.c_int, _p, _p]),
    "onda_append_update": (C.c_int, [_p, _p, _p, _p, C.c_int, C.c_int, _p]),
    "onda_prior_mix_stats": (C.c_int, [_p, _p, _p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int,
                                       _p, _p, _p, C.c_size_t, _p]),
    "onda_prior_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "onda_step_log_workspace_bytes": (C.c_size_t, []),
    "onda_step_log_stats": (C.c_int, [_p, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p, C.c_size_t, _p]),
    "onda_target_loss_workspace_bytes": (C.c_size_t, []),
    "onda_target_loss_fused": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, _p, C.c_float, C.c_float, C.c_float, C.c_int,
                                         _p, _p, _p, C.c_size_t, _p]),
    "onda_weight_ema_update": (C.c_int, [_p, C.c_int, C.c_float, C.c_float, _p]),
    "onda_confusion_update": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, C.c_int, _p, C.c_int, C.c_int, _p, _p, _p]),
    "onda_allreduce_oneshot": (C.c_int, [_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(_p), C.POINTER(_p),
                                         C.c_uint32, _p]),
    "onda_ema_update_and_table_allreduce": (C.c_int, [_p, _p, _p, _p, C.c_int, C.c_int, C.c_float, C.c_int, _p,
                                                      C.c_int, C.c_int, C.POINTER(_p), C.POINTER(_p), C.POINTER(_p),
                                                      C.c_uint32, _p, _p]),
}


class NativeError(RuntimeError):
    """A call into libonda_b200.so returned a non-zero status."""


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load the library once; raise if it has not been built (``python onda_b200/build.py``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(onda_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(_LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(rc: int, what: str = ""):
    if rc != OK:
        msg = load().onda_last_error()
        raise NativeError(f"{what or 'onda_b200'} failed ({rc}): {msg.decode() if msg else '?'}")


def ptr(t):
    """Device pointer of a tensor (or NULL for None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
