"""ctypes binding of libphynet_b200.so (C ABI in include/phynet_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is
present when a context is requested, this raises.  The library is built in-tree
by ``phynetpy_b200.build`` / ``__graft_entry__.build()``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# PNB_LIB selects a tuning variant of the same library (tools/variants.py); the product path is the in-tree build
LIB_PATH = os.environ.get("PNB_LIB") or os.path.join(HERE, "libphynet_b200.so")

PNB_OK = 0
PNB_ERR_INVALID = 1
PNB_ERR_CUDA = 2
PNB_ERR_NOMEM = 3
PNB_ERR_STATE = 4
PNB_ERR_UNSUPPORTED = 5

EVAL_FULL = 0x1
EVAL_PENDING = 0x2
EVAL_NOSTORE = 0x4
EVAL_OUT_DEVICE = 0x8
EVAL_RESCALE = 0x10
BATCH_BLEN_DEVICE = 0x20

# every symbol include/phynet_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = (
    "pnb_abi_version", "pnb_last_error", "pnb_device_count",
    "pnb_ctx_create", "pnb_ctx_destroy", "pnb_ctx_set_stream", "pnb_ctx_synchronize", "pnb_ctx_last_stats",
    "pnb_ctx_set_profiling", "pnb_ctx_last_timings",
    "pnb_model_create_eigen", "pnb_model_create_jc69", "pnb_model_create_explicit", "pnb_model_destroy",
    "pnb_pmatrix_batch", "pnb_compress",
    "pnb_locus_create", "pnb_locus_destroy", "pnb_locus_invalidate", "pnb_locus_read_partial",
    "pnb_eval", "pnb_commit", "pnb_rollback",
    "pnb_batch_create", "pnb_batch_destroy", "pnb_batch_info", "pnb_batch_traffic", "pnb_batch_blen_host", "pnb_batch_eval", "pnb_batch_set_peer_outputs", "pnb_batch_set_peer_barrier", "pnb_batch_barrier_timed_out",
    "pnb_ctx_create_host", "pnb_plan",
    "pnb_msnc_net_create", "pnb_msnc_net_destroy", "pnb_msnc_net_info", "pnb_msnc_net_set_max_frontier", "pnb_msnc_net_set_branch_thetas",
    "pnb_msnc_eval", "pnb_eval_move",
    "pnb_candidates_count", "pnb_candidates_fill",
    "pnb_simulate_alignments",
)


class PhyNetB200Error(RuntimeError):
    """A non-zero status from the C ABI (CUDA failure, state error, unsupported shape)."""

    def __init__(self, code: int, message: str):
        super().__init__("[pnb status %d] %s" % (code, message))
        self.code = code


_lib = None
_lib_lock = threading.Lock()

_vp = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_dbl_p = C.POINTER(C.c_double)


def load() -> C.CDLL:
    """Load the shared library (once); raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libphynet_b200.so is missing at %s -- build it with `python -m phynetpy_b200.build` "
                "(there is no CPU fallback for this path)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.pnb_abi_version.restype = C.c_int
        lib.pnb_last_error.restype = C.c_char_p
        sigs = {
            "pnb_device_count": [C.POINTER(C.c_int)],
            "pnb_ctx_create": [C.c_int, C.POINTER(_vp)],
            "pnb_ctx_destroy": [_vp],
            "pnb_ctx_set_stream": [_vp, _vp],
            "pnb_ctx_synchronize": [_vp],
            "pnb_ctx_last_stats": [_vp, C.POINTER(_i64)],
            "pnb_ctx_set_profiling": [_vp, C.c_int],
            "pnb_ctx_last_timings": [_vp, _dbl_p],
            "pnb_model_create_eigen": [_vp, _vp, _vp, _vp, _vp, C.POINTER(_vp)],
            "pnb_model_create_jc69": [_vp, C.POINTER(_vp)],
            "pnb_model_create_explicit": [_vp, _vp, C.POINTER(_vp)],
            "pnb_model_destroy": [_vp],
            "pnb_pmatrix_batch": [_vp, _vp, _vp, _i64, _vp],
            "pnb_compress": [_vp, C.c_int, _vp, _i32, _i64, _vp, C.POINTER(_i64), _vp, _i64, _vp, _vp, _vp],
            "pnb_locus_create": [_vp, _i32, _i64, _vp, _i64, _vp, C.POINTER(_vp)],
            "pnb_locus_destroy": [_vp],
            "pnb_locus_invalidate": [_vp],
            "pnb_locus_read_partial": [_vp, _i32, _vp],
            "pnb_eval": [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_uint32, _vp],
            "pnb_batch_create": [_vp, _vp, _i64, _vp, _vp, _vp, _vp, C.c_uint32, C.POINTER(_vp)],
            "pnb_batch_destroy": [_vp],
            "pnb_batch_info": [_vp, C.POINTER(_i64)],
            "pnb_batch_traffic": [_vp, C.POINTER(_i64)],
            "pnb_batch_blen_host": [_vp, C.POINTER(_vp)],
            "pnb_batch_eval": [_vp, _vp, _vp, C.c_double, C.c_uint32, _vp],
            "pnb_batch_set_peer_outputs": [_vp, _i32, _vp],
            "pnb_batch_set_peer_barrier": [_vp, _i32, C.POINTER(_vp), _vp, C.POINTER(_i32), C.c_uint64],
            "pnb_batch_barrier_timed_out": [_vp, C.POINTER(_i32)],
            "pnb_commit": [_vp, _i64, _vp],
            "pnb_rollback": [_vp, _i64, _vp],
            "pnb_ctx_create_host": [C.POINTER(_vp)],
            "pnb_plan": [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_uint32, _vp, _vp, _vp],
            "pnb_msnc_net_create": [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, C.POINTER(_vp)],
            "pnb_msnc_net_destroy": [_vp],
            "pnb_msnc_net_info": [_vp, C.POINTER(_i32)],
            "pnb_msnc_net_set_max_frontier": [_vp, _i64],
            "pnb_msnc_net_set_branch_thetas": [_vp, _vp, C.c_double],
            "pnb_msnc_eval": [_vp, _vp, _i64, _vp, _vp, _vp, _vp, C.c_double, _vp],
            "pnb_eval_move": [_vp, _vp, _vp, _i32, _vp, _vp, _vp, C.c_double, C.c_uint32, _vp, _vp, C.c_double, _vp],
            "pnb_simulate_alignments": [_vp, _i64, _vp, _vp, _vp, _vp, _i64, C.c_uint64, _vp, _vp, _vp],
            "pnb_candidates_count": [_vp, _i64, _vp, _vp, _vp, _i32, _vp, _vp],
            "pnb_candidates_fill": [_vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
        }
        for name, argtypes in sigs.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        if lib.pnb_abi_version() != 1:
            raise ImportError("libphynet_b200.so has ABI version %d, expected 1" % lib.pnb_abi_version())
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc == PNB_OK:
        return
    msg = load().pnb_last_error().decode("utf-8", "replace")
    if rc == PNB_ERR_INVALID:
        raise ValueError(msg)
    if rc == PNB_ERR_NOMEM:
        raise MemoryError(msg)
    raise PhyNetB200Error(rc, msg)


def ptr(a: Optional[np.ndarray]) -> Optional[int]:
    """Address of a C-contiguous NumPy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data
