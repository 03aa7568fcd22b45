"""Device context and batch evaluation helpers above the C ABI.

One :class:`DeviceContext` per process per GPU.  ``run_parallel_chains`` in the
reference pickles the sampler into spawned workers (reference
``src/_mcmc_seq.py:4144-4175``); device handles are therefore never pickled and
are re-created lazily in whichever process touches them (contexts are keyed by
PID).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .tree import FlatTree

_contexts: Dict[Tuple[int, int], "DeviceContext"] = {}

STAT_NAMES = ("updates", "ops", "tiles", "launches", "h2d_bytes", "d2h_bytes", "jobs_cached", "nodes_recomputed")


def default_device() -> int:
    for var in ("PHYNET_B200_DEVICE", "LOCAL_RANK"):
        v = os.environ.get(var)
        if v not in (None, ""):
            return int(v)
    return 0


class DeviceContext:
    """Owns a ``pnb_ctx`` (stream, staging buffers, device arena) on one GPU."""

    def __init__(self, device: Optional[int] = None):
        lib = _lib.load()
        self.device = default_device() if device is None else int(device)
        h = C.c_void_p()
        _lib.check(lib.pnb_ctx_create(self.device, C.byref(h)))
        self._h = h
        self._pid = os.getpid()
        self._lib = lib

    @property
    def handle(self) -> C.c_void_p:
        if self._h is None:
            raise RuntimeError("DeviceContext was destroyed")
        return self._h

    def set_stream(self, cuda_stream: Optional[int]) -> None:
        """Run this context's kernels on an existing CUDA stream (e.g. torch's current stream)."""
        _lib.check(self._lib.pnb_ctx_set_stream(self.handle, C.c_void_p(cuda_stream or 0)))

    def synchronize(self) -> None:
        _lib.check(self._lib.pnb_ctx_synchronize(self.handle))

    def last_stats(self) -> Dict[str, int]:
        arr = (C.c_int64 * 8)()
        _lib.check(self._lib.pnb_ctx_last_stats(self.handle, arr))
        return dict(zip(STAT_NAMES, (int(x) for x in arr)))

    def set_profiling(self, enable: bool) -> None:
        _lib.check(self._lib.pnb_ctx_set_profiling(self.handle, 1 if enable else 0))

    def last_timings(self) -> Dict[str, float]:
        """ms of the last evaluation: host compile, K-a kernel, pruning kernel, whole host call."""
        arr = (C.c_double * 4)()
        _lib.check(self._lib.pnb_ctx_last_timings(self.handle, arr))
        return {"compile_ms": arr[0], "pmatrix_ms": arr[1], "prune_ms": arr[2], "call_ms": arr[3]}

    def destroy(self) -> None:
        if self._h is not None and self._pid == os.getpid():
            self._lib.pnb_ctx_destroy(self._h)
        self._h = None

    # -- K-a ------------------------------------------------------------------
    def pmatrix_batch(self, model_handle: C.c_void_p, t: np.ndarray) -> np.ndarray:
        t = np.ascontiguousarray(t, dtype=np.float64).ravel()
        out = np.empty((t.shape[0], 4, 4), dtype=np.float64)
        _lib.check(self._lib.pnb_pmatrix_batch(self.handle, model_handle, _lib.ptr(t), t.shape[0], _lib.ptr(out)))
        return out

    # -- K-b / K-c ---------------------------------------------------------------
    def eval(self, model_handle: C.c_void_p, loci: np.ndarray, node_off: np.ndarray, parent: np.ndarray,
             blen: np.ndarray, tip_row: np.ndarray, site_rate: float = 1.0, flags: int = 0,
             P_edge: Optional[np.ndarray] = None, out_device_ptr: Optional[int] = None) -> Optional[np.ndarray]:
        """One ``pnb_eval`` over M (locus, tree) jobs given as CSR-packed arrays."""
        M = int(loci.shape[0])
        assert loci.dtype == np.uint64 and node_off.dtype == np.int64 and parent.dtype == np.int32
        assert blen.dtype == np.float64 and tip_row.dtype == np.int32
        if out_device_ptr is not None:
            flags |= _lib.EVAL_OUT_DEVICE
            out = None
            out_ptr = C.c_void_p(out_device_ptr)
        else:
            out = np.empty(M, dtype=np.float64)
            out_ptr = _lib.ptr(out)
        _lib.check(self._lib.pnb_eval(self.handle, model_handle, M, _lib.ptr(loci), _lib.ptr(node_off),
                                      _lib.ptr(parent), _lib.ptr(blen), _lib.ptr(tip_row),
                                      _lib.ptr(P_edge) if P_edge is not None else None,
                                      float(site_rate), flags, out_ptr))
        return out

    def eval_move(self, model_handle: C.c_void_p, locus_handle: C.c_void_p, parent: np.ndarray, blen: np.ndarray,
                  tip_row: np.ndarray, site_rate: float, flags: int, net_handle: C.c_void_p, node_species: np.ndarray,
                  theta: float) -> Tuple[float, float]:
        """Both likelihood factors of ONE locus' gene tree behind one host synchronisation (``pnb_eval_move``)."""
        assert parent.dtype == np.int32 and blen.dtype == np.float64 and tip_row.dtype == np.int32 and node_species.dtype == np.int32
        out = (C.c_double * 2)()
        _lib.check(self._lib.pnb_eval_move(self.handle, model_handle, locus_handle, int(parent.shape[0]), _lib.ptr(parent),
                                           _lib.ptr(blen), _lib.ptr(tip_row), float(site_rate), flags, net_handle,
                                           _lib.ptr(node_species), float(theta), out))
        return out[0], out[1]

    def commit(self, loci: np.ndarray) -> None:
        _lib.check(self._lib.pnb_commit(self.handle, int(loci.shape[0]), _lib.ptr(loci)))

    def rollback(self, loci: np.ndarray) -> None:
        _lib.check(self._lib.pnb_rollback(self.handle, int(loci.shape[0]), _lib.ptr(loci)))


class ResidentBatch:
    """A fixed set of (locus, topology) jobs resident in HBM (``pnb_batch_*``): evaluations send branch
    lengths only.  ``flags``: 0 (every evaluation is committed to the loci's partial caches) or
    ``_lib.EVAL_NOSTORE``, optionally ``| _lib.EVAL_RESCALE``."""

    def __init__(self, ctx: DeviceContext, model_handle: C.c_void_p, loci: np.ndarray, node_off: np.ndarray,
                 parent: np.ndarray, tip_row: Optional[np.ndarray], flags: int = 0):
        assert loci.dtype == np.uint64 and node_off.dtype == np.int64 and parent.dtype == np.int32
        assert tip_row is None or tip_row.dtype == np.int32
        self._ctx = ctx
        self._lib = ctx._lib
        h = C.c_void_p()
        _lib.check(self._lib.pnb_batch_create(ctx.handle, model_handle, int(loci.shape[0]), _lib.ptr(loci), _lib.ptr(node_off),
                                              _lib.ptr(parent), _lib.ptr(tip_row), flags, C.byref(h)))
        self._h = h
        self._pid = os.getpid()
        self.n_jobs = int(loci.shape[0])
        arr = (C.c_int64 * 8)()
        _lib.check(self._lib.pnb_batch_info(h, arr))
        self.info = dict(zip(("jobs", "nodes", "ops", "tiles", "updates", "internal_nodes", "smem_bytes", "chunks"),
                             (int(x) for x in arr)))
        self.n_nodes = self.info["nodes"]
        _lib.check(self._lib.pnb_batch_traffic(h, arr))
        self.traffic = {"partials_written": int(arr[0]), "partials_reread": int(arr[1]), "tip_codes": int(arr[2]),
                        "weights_programs_pmatrices": int(arr[3])}
        self._pinned: Optional[np.ndarray] = None

    @property
    def blen_host(self) -> np.ndarray:
        """The batch's pinned staging buffer as a float64 array (``n_nodes``): lengths written here upload without a copy."""
        if self._pinned is None:
            p = C.c_void_p()
            _lib.check(self._lib.pnb_batch_blen_host(self._h, C.byref(p)))
            buf = (C.c_double * self.n_nodes).from_address(p.value)
            self._pinned = np.frombuffer(buf, dtype=np.float64)
        return self._pinned

    def eval(self, model_handle: C.c_void_p, blen, site_rate: float = 1.0, out_device_ptr: Optional[int] = None
             ) -> Optional[np.ndarray]:
        """``blen``: float64 host array in the node order of the creation arrays, or an ``int`` device address."""
        flags = 0
        if isinstance(blen, (int, np.integer)):
            flags |= _lib.BATCH_BLEN_DEVICE
            blen_ptr = C.c_void_p(int(blen))
        else:
            assert blen.dtype == np.float64 and blen.shape[0] == self.n_nodes and blen.flags["C_CONTIGUOUS"]
            blen_ptr = _lib.ptr(blen)
        if out_device_ptr is not None:
            flags |= _lib.EVAL_OUT_DEVICE
            out, out_ptr = None, C.c_void_p(out_device_ptr)
        else:
            out = np.empty(self.n_jobs, dtype=np.float64)
            out_ptr = _lib.ptr(out)
        _lib.check(self._lib.pnb_batch_eval(self._h, model_handle, blen_ptr, float(site_rate), flags, out_ptr))
        return out

    def set_peer_outputs(self, peer_ptrs: Sequence[int]) -> None:
        """Fused collective: device addresses (valid in this process) of the same output slot in every PEER's
        vector; evaluations with ``out_device_ptr`` then store each job's logL there too.  ``[]`` turns it off."""
        arr = (C.c_void_p * max(len(peer_ptrs), 1))(*[C.c_void_p(int(p)) for p in peer_ptrs])
        _lib.check(self._lib.pnb_batch_set_peer_outputs(self._h, len(peer_ptrs), arr))

    def set_peer_barrier(self, peer_flag_ptrs: Sequence[int], my_flags_ptr: int, peer_ranks: Sequence[int], epoch: int) -> None:
        """Arm the kernel-issued barrier of the fused collective for the NEXT evaluation (``ShardPlan.peer_barrier_args``
        gives the three address arguments): the block that finishes the evaluation's last job publishes ``epoch`` to every
        peer and waits for theirs, so no barrier launch follows the kernel.  ``epoch`` = 1, 2, 3, ... in the order all
        ranks run their sharded evaluations."""
        n = len(peer_flag_ptrs)
        arr = (C.c_void_p * max(n, 1))(*[C.c_void_p(int(p)) for p in peer_flag_ptrs])
        ranks = (C.c_int32 * max(n, 1))(*[int(r) for r in peer_ranks])
        _lib.check(self._lib.pnb_batch_set_peer_barrier(self._h, n, arr, C.c_void_p(int(my_flags_ptr)), ranks, C.c_uint64(int(epoch))))

    def barrier_timed_out(self) -> bool:
        v = C.c_int32(0)
        _lib.check(self._lib.pnb_batch_barrier_timed_out(self._h, C.byref(v)))
        return bool(v.value)

    def close(self) -> None:
        if self._h is not None and self._pid == os.getpid() and self._ctx._h is not None:
            self._pinned = None
            self._lib.pnb_batch_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_context(device: Optional[int] = None) -> DeviceContext:
    dev = default_device() if device is None else int(device)
    key = (os.getpid(), dev)
    ctx = _contexts.get(key)
    if ctx is None or ctx._h is None:
        ctx = DeviceContext(dev)
        _contexts[key] = ctx
    return ctx


def pack_trees(trees: Sequence[FlatTree], row_maps: Sequence[dict]) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Concatenate flat trees into the CSR arrays ``pnb_eval`` takes."""
    if len(trees) == 1:  # the single-locus MCMC step: no concatenation
        t = trees[0]
        return (np.array([0, t.n_nodes], dtype=np.int64), t.parent, t.blen, t.tip_rows(row_maps[0]))
    sizes = np.fromiter((t.n_nodes for t in trees), dtype=np.int64, count=len(trees))
    node_off = np.zeros(len(trees) + 1, dtype=np.int64)
    np.cumsum(sizes, out=node_off[1:])
    parent = np.concatenate([t.parent for t in trees]) if trees else np.zeros(0, np.int32)
    blen = np.concatenate([t.blen for t in trees]) if trees else np.zeros(0, np.float64)
    tip_row = np.concatenate([t.tip_rows(m) for t, m in zip(trees, row_maps)]) if trees else np.zeros(0, np.int32)
    return node_off, np.ascontiguousarray(parent, np.int32), np.ascontiguousarray(blen, np.float64), \
        np.ascontiguousarray(tip_row, np.int32)
