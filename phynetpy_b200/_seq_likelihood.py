"""Drop-in for the reference's ``phynetpy._seq_likelihood`` hot path, on B200.

Same public names, signatures, attributes and exceptions as the reference
(``src/_seq_likelihood.py``):

    tip_partials_for_char, DNA_STATES,
    SubstitutionModel(pi, rates), JC69(), HKY85(kappa, pi), GTR(pi, rates),
    FelsensteinCalculator(alignment).log_likelihood(gene_tree, model, *, site_rate=1.0)

but the arithmetic runs in hand-written sm_100a kernels behind the C ABI
(``include/phynet_b200.h``): site-pattern compression (K-d), batched P(t) (K-a)
and post-order pruning with the fused root reduction (K-b/K-c).  There is no CPU
fallback -- without the built library and a CUDA device these classes raise.

What stays on the host, deliberately: validation and the 4x4 eigendecomposition
(``np.linalg.eigh``, exactly the reference's call, so the eigenbasis is the same).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from .device import DeviceContext, default_context, pack_trees
from .tree import FlatTree, flatten

__all__ = [
    "SubstitutionModel", "JC69", "GTR", "HKY85", "FelsensteinCalculator",
    "DNA_STATES", "tip_partials_for_char",
]

# ======================================================================
# Nucleotide encoding (reference src/_seq_likelihood.py:104-140)
# ======================================================================
DNA_STATES: str = "ACGT"
_STATE_BIT = {"A": 1, "C": 2, "G": 4, "T": 8}
_AMBIGUITY = {
    "A": "A", "C": "C", "G": "G", "T": "T", "U": "T",
    "R": "AG", "Y": "CT", "S": "CG", "W": "AT", "K": "GT", "M": "AC",
    "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG",
    "N": "ACGT", "?": "ACGT", "-": "ACGT", "X": "ACGT", ".": "ACGT", "O": "ACGT",
}


def _mask_for_key(key: str) -> int:
    """4-bit tip mask (A=1,C=2,G=4,T=8) of an upper-cased column character;
    anything not in the IUPAC table is fully ambiguous (reference :134-137)."""
    allowed = _AMBIGUITY.get(key)
    if allowed is None:
        return 0xF
    m = 0
    for nt in allowed:
        m |= _STATE_BIT[nt]
    return m


# byte -> mask for the ASCII fast path (bytes are already upper-cased)
_CODE_OF_BYTE = np.array([_mask_for_key(chr(b).upper()) if len(chr(b).upper()) == 1 else 0xF for b in range(256)],
                         dtype=np.uint8)


def tip_partials_for_char(ch: str) -> np.ndarray:
    """Length-4 tip partial vector (order A,C,G,T) for a DNA character
    (same contract as reference ``tip_partials_for_char``, :120-140)."""
    m = _mask_for_key(str(ch).upper())
    return np.array([(m >> j) & 1 for j in range(4)], dtype=np.float64)


# ======================================================================
# Substitution models (reference :147-304)
# ======================================================================
class SubstitutionModel:
    """Time-reversible 4-state model; ``p_matrix(t)`` is evaluated on the GPU (K-a).

    Construction mirrors the reference (:162-227): Q[i,j] = r_ij * pi_j,
    rows sum to 0, normalised to unit mean rate, eigendecomposed through the
    pi^(1/2) similarity transform with ``np.linalg.eigh``.
    """

    def __init__(self, pi: Sequence[float], rates: Sequence[float]) -> None:
        self.pi = np.asarray(pi, dtype=np.float64)
        if self.pi.shape != (4,):
            raise ValueError("Base frequencies must have length 4 (A,C,G,T).")
        if not math.isclose(float(self.pi.sum()), 1.0, abs_tol=1e-9):
            raise ValueError("Base frequencies must sum to 1.")
        self._rates = np.asarray(rates, dtype=np.float64)
        self._build_and_decompose()

    def _build_Q(self) -> np.ndarray:
        ac, ag, at_, cg, ct, gt = self._rates  # order AC, AG, AT, CG, CT, GT (:189)
        r = np.array([[0.0, ac, ag, at_], [ac, 0.0, cg, ct], [ag, cg, 0.0, gt], [at_, ct, gt, 0.0]],
                     dtype=np.float64)
        return r * self.pi[np.newaxis, :]

    def _build_and_decompose(self) -> None:
        Q = self._build_Q()
        np.fill_diagonal(Q, 0.0)
        np.fill_diagonal(Q, -Q.sum(axis=1))
        mean_rate = -float(np.dot(self.pi, np.diag(Q)))
        if mean_rate <= 0.0:
            raise ValueError("Degenerate substitution model (non-positive rate).")
        Q /= mean_rate
        self.Q = Q
        sqrt_pi = np.sqrt(self.pi)
        inv_sqrt_pi = 1.0 / sqrt_pi
        B = (Q * sqrt_pi[:, None]) * inv_sqrt_pi[None, :]
        B = 0.5 * (B + B.T)
        evals, evecs = np.linalg.eigh(B)
        self._evals = evals
        self._U = inv_sqrt_pi[:, None] * evecs
        self._Uinv = evecs.T * sqrt_pi[None, :]

    # -- device side -------------------------------------------------------
    def p_matrix(self, t: float) -> np.ndarray:
        """``P(t) = exp(Q t)`` as a 4x4 row-stochastic matrix (negative t -> 0)."""
        return self.p_matrix_batch(np.array([t], dtype=np.float64))[0]

    def p_matrix_batch(self, ts: Sequence[float], context: Optional[DeviceContext] = None) -> np.ndarray:
        """P(t) for every t at once: ``(len(ts), 4, 4)``.  One K-a launch."""
        ctx = context or default_context()
        return ctx.pmatrix_batch(_model_handle(self, ctx), np.asarray(ts, dtype=np.float64))

    def __getstate__(self) -> Dict[str, Any]:
        d = dict(self.__dict__)
        d.pop("_pnb_handles", None)  # device handles never cross a process boundary
        return d


class JC69(SubstitutionModel):
    """Jukes-Cantor (1969); closed-form P(t) (reference :247-268), evaluated on the GPU."""

    def __init__(self) -> None:
        super().__init__([0.25, 0.25, 0.25, 0.25], [1.0] * 6)

    # Defined (not inherited) so that, as in the reference, ``JC69.p_matrix`` is a
    # distinct method: the device uses the closed form 1/4 +- e^{-4t/3} for it.
    def p_matrix(self, t: float) -> np.ndarray:
        return self.p_matrix_batch(np.array([t], dtype=np.float64))[0]


class HKY85(SubstitutionModel):
    """Hasegawa-Kishino-Yano (1985): rates [1,kappa,1,1,kappa,1] (reference :271-286)."""

    def __init__(self, kappa: float, pi: Sequence[float]) -> None:
        self.kappa = float(kappa)
        super().__init__(pi, [1.0, self.kappa, 1.0, 1.0, self.kappa, 1.0])


class GTR(SubstitutionModel):
    """General time-reversible model (reference :291-304)."""

    def __init__(self, pi: Sequence[float], rates: Sequence[float]) -> None:
        super().__init__(pi, rates)


def _device_kind(model: Any) -> str:
    """eigen | jc69 | explicit -- how the device obtains P(t) for this model object.

    A subclass that overrides ``p_matrix`` itself (as reference users may) is
    honoured: its matrices are computed by that method on the host and uploaded.
    Model objects of the reference package are recognised by method identity.
    """
    pm = getattr(type(model), "p_matrix", None)
    if isinstance(model, SubstitutionModel):
        if pm is SubstitutionModel.p_matrix:
            return "eigen"
        if pm is JC69.p_matrix:
            return "jc69"
        return "explicit"
    qn = getattr(pm, "__qualname__", "")
    if qn == "SubstitutionModel.p_matrix" and all(hasattr(model, a) for a in ("_evals", "_U", "_Uinv", "pi")):
        return "eigen"
    if qn == "JC69.p_matrix":
        return "jc69"
    return "explicit"


class _ModelHandle:
    __slots__ = ("handle", "fingerprint", "lib", "pid")

    def __init__(self, handle, fingerprint, lib):
        self.handle = handle
        self.fingerprint = fingerprint
        self.lib = lib
        self.pid = os.getpid()

    def __del__(self):
        try:
            if self.handle is not None and self.pid == os.getpid():
                self.lib.pnb_model_destroy(self.handle)
        except Exception:
            pass


# handles of model objects that cannot carry an attribute (``__slots__`` classes): id(model) -> (weakref | model, cache)
_slot_model_handles: Dict[int, Any] = {}


def _handle_cache(model: Any) -> Dict[Any, "_ModelHandle"]:
    """Where the device handles of a model object live: on the object when it has a ``__dict__`` (they go away
    with it and are dropped by ``__getstate__``), else in a module-level table keyed by identity that holds the
    object weakly when it can (the entry is dropped when the model dies) and strongly when it cannot."""
    d = getattr(model, "__dict__", None)
    if d is not None:
        cache = d.get("_pnb_handles")
        if cache is None:
            cache = d["_pnb_handles"] = {}
        return cache
    import weakref
    key = id(model)
    hit = _slot_model_handles.get(key)
    if hit is not None:
        owner = hit[0]() if isinstance(hit[0], weakref.ref) else hit[0]
        if owner is model:
            return hit[1]
    cache: Dict[Any, "_ModelHandle"] = {}
    try:
        ref: Any = weakref.ref(model, lambda _r, k=key: _slot_model_handles.pop(k, None))
    except TypeError:
        ref = model
    _slot_model_handles[key] = (ref, cache)
    return cache


def _model_handle(model: Any, ctx: DeviceContext) -> C.c_void_p:
    """``pnb_model*`` for (model, context); rebuilt when the parameters change."""
    kind = _device_kind(model)
    pi = np.ascontiguousarray(model.pi, dtype=np.float64)
    if kind == "eigen":
        ev = np.ascontiguousarray(model._evals, dtype=np.float64)
        U = np.ascontiguousarray(model._U, dtype=np.float64)
        Ui = np.ascontiguousarray(model._Uinv, dtype=np.float64)
        fp = (kind, pi.tobytes(), ev.tobytes(), U.tobytes(), Ui.tobytes())
    elif kind == "explicit":
        # user code computes P(t): the only way to notice that its parameters changed is to ask it.  Two probe
        # matrices are part of the identity, so cached partials of the old parameters are never reused.
        probe = b"".join(np.ascontiguousarray(model.p_matrix(t), dtype=np.float64).tobytes() for t in (0.0625, 1.0))
        fp = (kind, pi.tobytes(), probe)
    else:
        fp = (kind, pi.tobytes())
    cache = _handle_cache(model)
    key = (os.getpid(), id(ctx))
    mh = cache.get(key)
    if mh is not None and mh.fingerprint == fp:
        return mh.handle
    lib = _lib.load()
    h = C.c_void_p()
    if kind == "eigen":
        _lib.check(lib.pnb_model_create_eigen(ctx.handle, _lib.ptr(ev), _lib.ptr(U), _lib.ptr(Ui), _lib.ptr(pi), C.byref(h)))
    elif kind == "jc69":
        _lib.check(lib.pnb_model_create_jc69(ctx.handle, C.byref(h)))
    else:
        _lib.check(lib.pnb_model_create_explicit(ctx.handle, _lib.ptr(pi), C.byref(h)))
    cache[key] = _ModelHandle(h, fp, lib)
    return h


def _explicit_P_edges(model: Any, blen: np.ndarray, site_rate: float) -> np.ndarray:
    """Host P matrices for a model whose ``p_matrix`` is user code (reference :419-422)."""
    out = np.empty((blen.shape[0], 16), dtype=np.float64)
    for i, b in enumerate(blen):
        t = 0.0 if math.isnan(b) else float(b) * site_rate
        out[i] = np.asarray(model.p_matrix(t), dtype=np.float64).reshape(16)
    return out


# ======================================================================
# Alignment -> bytes (reference :338-349 keys on ch.upper())
# ======================================================================
def _alignment_matrix(seqs: List[str]):
    """(n, L) uint8 matrix of upper-cased characters + the byte -> mask table.

    ASCII sequences map to their own bytes.  Anything else (characters whose
    ``upper()`` is not a single latin-1 byte) goes through a per-locus table so
    that distinct upper-cased keys stay distinct patterns, as in the reference.
    """
    n = len(seqs)
    L = len(seqs[0])
    if all(s.isascii() for s in seqs):
        mat = np.empty((n, L), dtype=np.uint8)
        for r, s in enumerate(seqs):
            mat[r] = np.frombuffer(s.upper().encode("ascii"), dtype=np.uint8)
        return mat, _CODE_OF_BYTE
    ids: Dict[str, int] = {}
    lut = np.full(256, 0xF, dtype=np.uint8)
    mat = np.empty((n, L), dtype=np.uint8)
    for r, s in enumerate(seqs):
        row = mat[r]
        for c, ch in enumerate(s):
            key = ch.upper()
            i = ids.get(key)
            if i is None:
                i = len(ids)
                if i >= 256:
                    raise ValueError("More than 256 distinct characters in one locus.")
                ids[key] = i
                lut[i] = _mask_for_key(key)
            row[c] = i
    return mat, lut


_DEVICE_COMPRESS_MIN_BYTES = 1 << 22  # alignments of >= 4 MiB are compressed on the GPU


def compress_alignment(mat: np.ndarray, lut: np.ndarray, context: Optional[DeviceContext] = None,
                       use_device: Optional[bool] = None):
    """K-d through the C ABI: (n, L) bytes -> codes (n, P), int weights (P), first_col (P), pattern_of_col (L)."""
    lib = _lib.load()
    n, L = mat.shape
    mat = np.ascontiguousarray(mat, dtype=np.uint8)
    if L == 0:
        return (np.zeros((n, 0), np.uint8), np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int32))
    if use_device is None:
        use_device = mat.size >= _DEVICE_COMPRESS_MIN_BYTES
    ctx_h = None
    if use_device:
        ctx_h = (context or default_context()).handle
    codes = np.empty((n, L), dtype=np.uint8)
    weights = np.empty(L, dtype=np.int64)
    first = np.empty(L, dtype=np.int64)
    poc = np.empty(L, dtype=np.int32)
    P = C.c_int64(0)
    _lib.check(lib.pnb_compress(ctx_h, 1 if use_device else 0, _lib.ptr(mat), n, L, _lib.ptr(np.ascontiguousarray(lut)),
                                C.byref(P), _lib.ptr(codes), L, _lib.ptr(weights), _lib.ptr(first), _lib.ptr(poc)))
    p = int(P.value)
    return np.ascontiguousarray(codes[:, :p]), weights[:p].copy(), first[:p].copy(), poc


# ======================================================================
# Felsenstein pruning (reference :311-426)
# ======================================================================
class FelsensteinCalculator:
    """Per-locus likelihood ``P(S | g)``; patterns, tip codes and per-node partials live in HBM.

    Attributes (as in the reference): ``taxa`` -- leaf labels in alignment order;
    ``n_patterns`` -- number of distinct site patterns.

    Beyond the reference: the partials of the last evaluated tree stay cached on
    the device, so re-scoring after a local change (one branch length, one NNI)
    recomputes only the nodes on the path to the root.  Results are identical to
    a full recomputation; ``invalidate()`` drops the cache.
    """

    def __init__(self, alignment: Dict[str, str], *, context: Optional[DeviceContext] = None,
                 device_compress: Optional[bool] = None) -> None:
        if not alignment:
            raise ValueError("Empty alignment passed to FelsensteinCalculator.")
        self.taxa: List[str] = list(alignment.keys())
        seqs = [alignment[t] for t in self.taxa]
        length = len(seqs[0])
        if any(len(s) != length for s in seqs):
            raise ValueError("All sequences in a locus must be equally long.")
        mat, lut = _alignment_matrix(seqs)
        codes, iw, first, _ = compress_alignment(mat, lut, context, device_compress)
        self._init_from_codes(codes, iw, context)
        self._first_col = first

    @classmethod
    def from_file(cls, path, format: Optional[str] = None, **kwargs) -> "FelsensteinCalculator":
        """Build from a FASTA / NEXUS / PHYLIP file (``ingest.read_alignment``; the extension decides,
        NEXUS by default as in the reference's ``_read_alignment``, ``src/data/_sequence.py:258-300``)."""
        from .ingest import read_alignment
        return cls(read_alignment(path, format), **kwargs)

    @classmethod
    def from_matrix(cls, taxa: Sequence[str], mat: np.ndarray, *, context: Optional[DeviceContext] = None,
                    device_compress: Optional[bool] = None) -> "FelsensteinCalculator":
        """Build from an (n_taxa, L) uint8 matrix of upper-case ASCII characters (no
        Python strings; large synthetic alignments)."""
        mat = np.ascontiguousarray(mat, dtype=np.uint8)
        if mat.ndim != 2 or mat.shape[0] != len(taxa) or mat.shape[0] == 0:
            raise ValueError("mat must be (n_taxa, L) with n_taxa >= 1")
        self = cls.__new__(cls)
        self.taxa = list(taxa)
        codes, iw, first, _ = compress_alignment(mat, _CODE_OF_BYTE, context, device_compress)
        self._init_from_codes(codes, iw, context)
        self._first_col = first
        return self

    @classmethod
    def from_codes(cls, taxa: Sequence[str], codes: np.ndarray, weights: np.ndarray,
                   *, context: Optional[DeviceContext] = None) -> "FelsensteinCalculator":
        """Build from already-compressed data: (n_taxa, P) 4-bit masks + integer weights."""
        self = cls.__new__(cls)
        self.taxa = list(taxa)
        self._init_from_codes(np.ascontiguousarray(codes, dtype=np.uint8), np.asarray(weights, dtype=np.int64), context)
        self._first_col = None
        return self

    def _init_from_codes(self, codes: np.ndarray, int_weights: np.ndarray, context: Optional[DeviceContext]) -> None:
        if codes.ndim != 2 or codes.shape[0] != len(self.taxa) or codes.shape[1] != int_weights.shape[0]:
            raise ValueError("codes must be (n_taxa, n_patterns) and weights (n_patterns,)")
        self.n_patterns: int = int(codes.shape[1])
        self._codes = codes
        self._int_weights = np.ascontiguousarray(int_weights, dtype=np.int64)
        self._weights = self._int_weights.astype(np.float64)  # reference keeps float64 weights (:359)
        self._row_of = {t: i for i, t in enumerate(self.taxa)}
        self._ctx = context
        self._locus: Optional[C.c_void_p] = None
        self._locus_pid = -1
        self._handle_arr: Optional[np.ndarray] = None

    # -- compatibility view of the reference's private table (:361-367) ------
    @property
    def _tip_partials(self) -> Dict[str, np.ndarray]:
        bits = np.array([[(c >> j) & 1 for j in range(4)] for c in range(16)], dtype=np.float64)
        return {t: bits[self._codes[r]] for r, t in enumerate(self.taxa)}

    # -- device residency ---------------------------------------------------------
    def _context(self) -> DeviceContext:
        if self._ctx is None or self._ctx._h is None or self._ctx._pid != os.getpid():
            self._ctx = default_context()
        return self._ctx

    def _locus_handle(self) -> C.c_void_p:
        if self._locus is not None and self._locus_pid == os.getpid():
            return self._locus
        ctx = self._context()
        lib = _lib.load()
        h = C.c_void_p()
        ld = max(self.n_patterns, 1)
        codes = self._codes if self.n_patterns else np.zeros((len(self.taxa), 1), np.uint8)
        _lib.check(lib.pnb_locus_create(ctx.handle, len(self.taxa), self.n_patterns, _lib.ptr(np.ascontiguousarray(codes)),
                                        ld, _lib.ptr(self._int_weights) if self.n_patterns else None, C.byref(h)))
        self._locus = h
        self._locus_pid = os.getpid()
        self._handle_arr = np.array([h.value], dtype=np.uint64)
        return h

    def invalidate(self) -> None:
        """Drop the cached per-node partials (the next evaluation recomputes every node)."""
        if self._locus is not None and self._locus_pid == os.getpid():
            _lib.check(_lib.load().pnb_locus_invalidate(self._locus))

    def close(self) -> None:
        if self._locus is not None and self._locus_pid == os.getpid() and self._ctx is not None and self._ctx._h is not None:
            _lib.load().pnb_locus_destroy(self._locus)
        self._locus = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __getstate__(self) -> Dict[str, Any]:
        d = dict(self.__dict__)
        d["_locus"] = None
        d["_ctx"] = None
        d["_handle_arr"] = None
        d["_locus_pid"] = -1
        return d

    # -- evaluation ---------------------------------------------------------------
    def log_likelihood(self, gene_tree: Any, model: Any, *, site_rate: float = 1.0,
                       full: bool = False, rescale: bool = False) -> float:
        """Felsenstein log-likelihood of the alignment on ``gene_tree`` (reference :369-397).

        ``gene_tree`` is borrowed and not mutated; any object with the reference's
        ``Network`` read API, a :class:`~phynetpy_b200.tree.GeneTree` or a
        :class:`~phynetpy_b200.tree.FlatTree`.  ``full=True`` ignores cached partials.

        ``rescale=True`` is an option beyond the reference: every node is renormalised by an
        exact power of two and carries its exponent, so trees of thousands of taxa do not
        underflow.  The default reproduces the reference exactly, including its floor of
        ``1e-300`` per site (:396) when a site likelihood underflows.
        """
        ft = flatten(gene_tree)
        ctx = self._context()
        self._locus_handle()
        node_off, parent, blen, tip_row = pack_trees([ft], [self._row_of])
        flags = (_lib.EVAL_FULL if full else 0) | (_lib.EVAL_RESCALE if rescale else 0)
        mh = _model_handle(model, ctx)
        P_edge = _explicit_P_edges(model, blen, float(site_rate)) if _device_kind(model) == "explicit" else None
        out = ctx.eval(mh, self._handle_arr, node_off, parent, blen, tip_row, float(site_rate), flags, P_edge)
        return float(out[0])

    def node_partial(self, node_index: int) -> np.ndarray:
        """(n_patterns, 4) partial of a node of the last evaluated tree (flat index); for tests."""
        out = np.empty((self.n_patterns, 4), dtype=np.float64)
        _lib.check(_lib.load().pnb_locus_read_partial(self._locus_handle(), int(node_index), _lib.ptr(out)))
        return out
