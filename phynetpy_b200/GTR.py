"""Public time-reversible substitution models with ``expt(t) = e^{Qt}`` on the GPU.

Mirror of the reference's ``phynetpy.GTR`` module surface (``src/GTR.py``):

    JC()  K80(kappa)  F81(base_freqs)  HKY(base_freqs, kappa)
    TN93(base_freqs, kappa_r, kappa_y)  K81(alpha, beta, gamma)
    SYM(transitions)  GTR(base_freqs, transitions)

with ``getQ``, ``expt``, ``set_hyperparams`` / ``get_hyperparams``, ``state_count``,
``SubstitutionModelError`` and the per-class ``HYPERPARAMS``.  Rate order is
AC, AG, AT, CG, CT, GT (``src/GTR.py:473-516``).

Design difference: the reference evaluates six closed forms plus a cached
eigendecomposition, one 4x4 at a time; here every class hands its (normalised) Q's
eigen-system to ONE batched device kernel (K-a, ``pnb_pmatrix_batch``), because
``e^{Qt} = U diag(e^{lambda t}) U^{-1}`` is the same function as each closed form
(the reference's own tests pin them together at 1e-11, ``tests/test_gtr.py:221-232,
498-533``) and a branch batch is what the pruning path needs.  ``expt_batch(ts)``
is the efficient entry point; ``expt(t)`` is the reference-shaped one.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from ._seq_likelihood import _ModelHandle  # noqa: F401  (handle lifetime helper)
from . import _lib
from .device import DeviceContext, default_context

import ctypes as C
import os

_SUM_TOL = 1e-9


class SubstitutionModelError(Exception):
    """Malformed substitution-model inputs (reference ``src/GTR.py:110-130``)."""

    def __init__(self, message: str = "Unknown substitution model error") -> None:
        self.message = message
        super().__init__(self.message)


def _as_float_list(values: Any) -> List[float]:
    """Any array-like of numbers as a FLAT list of floats: a column vector such as ``np.ones((6, 1))`` is
    flattened, as in the reference (``src/GTR.py:154-168``), instead of iterating as length-1 arrays."""
    return np.asarray(values, dtype=np.double).reshape(-1).tolist()


def _positive(name: str, value: Any) -> float:
    try:
        v = float(value)
    except (TypeError, ValueError) as exc:
        raise SubstitutionModelError("%s must be a number, got %r" % (name, value)) from exc
    if not math.isfinite(v) or v <= 0.0:
        raise SubstitutionModelError("%s must be positive and finite, got %r" % (name, value))
    return v


def _rate_index(i: int, j: int, states: int) -> int:
    """Upper-triangular row-major index of the unordered pair {i, j}."""
    if i > j:
        i, j = j, i
    return i * (2 * states - i - 1) // 2 + (j - i - 1)


class GTR:
    """Generalised time-reversible model; superclass of the named special cases."""

    HYPERPARAMS: Tuple[str, ...] = ("states", "base frequencies", "transitions")

    def __init__(self, base_freqs: Sequence[float], transitions: Sequence[float], states: int = 4) -> None:
        self.states: int = states
        self.freqs: List[float] = _as_float_list(base_freqs)
        self.trans: List[float] = _as_float_list(transitions)
        self._is_valid(self.trans, self.freqs, self.states)
        self._decomp: Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]] = None
        self._handle = None
        self.Q = self.buildQ()
        self.Qt: Optional[np.ndarray] = None

    # -- accessors ----------------------------------------------------------
    def getQ(self) -> np.ndarray:
        return self.Q

    def get_hyperparams(self) -> Tuple[List[float], List[float]]:
        return self.freqs, self.trans

    def state_count(self) -> int:
        return self.states

    def set_hyperparams(self, params: Dict[str, Any]) -> None:
        unknown = sorted(set(params) - set(self.HYPERPARAMS))
        if unknown:
            accepted = ", ".join(self.HYPERPARAMS) or "(none)"
            raise SubstitutionModelError("%s does not have hyperparameter(s) %s. Accepted: %s."
                                         % (type(self).__name__, unknown, accepted))
        self._apply_hyperparams(params)
        self._is_valid(self.trans, self.freqs, self.states)
        self.buildQ()

    def _apply_hyperparams(self, params: Dict[str, Any]) -> None:
        if "states" in params:
            self.states = int(params["states"])
        if "transitions" in params:
            self.trans = _as_float_list(params["transitions"])
        if "base frequencies" in params:
            self.freqs = _as_float_list(params["base frequencies"])

    # -- Q and its eigen-system (host; reference :473-557) ---------------------
    def buildQ(self) -> np.ndarray:
        freqs = np.asarray(self.freqs, dtype=np.double)
        rates = np.asarray(self.trans, dtype=np.double)
        n = self.states
        Q = np.zeros((n, n), dtype=np.double)
        for i in range(n):
            for j in range(n):
                if i != j:
                    Q[i][j] = rates[_rate_index(i, j, n)] * freqs[j]
        np.fill_diagonal(Q, -Q.sum(axis=1))
        mean_rate = -float(freqs @ np.diag(Q))
        if mean_rate <= 0.0:
            raise SubstitutionModelError("Degenerate substitution model: the overall substitution rate is not positive.")
        self.Q = Q / mean_rate
        self._decomp = None   # cache invalidation, as the reference does (:515)
        self._handle = None
        return self.Q

    def _decompose(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        if self._decomp is not None:
            return self._decomp
        freqs = np.asarray(self.freqs, dtype=np.double)
        if np.all(freqs > 0.0):
            sqrt_pi = np.sqrt(freqs)
            B = (self.Q * sqrt_pi[:, None]) / sqrt_pi[None, :]
            B = 0.5 * (B + B.T)
            evals, evecs = np.linalg.eigh(B)
            U = evecs / sqrt_pi[:, None]
            U_inv = evecs.T * sqrt_pi[None, :]
        else:
            # a zero base frequency breaks the similarity transform (reference :548-554)
            evals, U = np.linalg.eig(self.Q)
            evals = np.real(evals)
            U = np.real(U)
            U_inv = np.linalg.inv(U)
        self._decomp = (evals, U, U_inv)
        return self._decomp

    # -- device ------------------------------------------------------------------
    def _device_handle(self, ctx: DeviceContext):
        if self.states != 4:
            raise SubstitutionModelError("the device P(t) kernel is 4-state (DNA) only; got states=%d" % self.states)
        key = (os.getpid(), id(ctx))
        if self._handle is not None and self._handle[0] == key:
            return self._handle[1].handle
        evals, U, U_inv = self._decompose()
        pi = np.ascontiguousarray(self.freqs, dtype=np.float64)
        h = C.c_void_p()
        lib = _lib.load()
        _lib.check(lib.pnb_model_create_eigen(ctx.handle, _lib.ptr(np.ascontiguousarray(evals, dtype=np.float64)),
                                              _lib.ptr(np.ascontiguousarray(U, dtype=np.float64)),
                                              _lib.ptr(np.ascontiguousarray(U_inv, dtype=np.float64)),
                                              _lib.ptr(pi), C.byref(h)))
        self._handle = (key, _ModelHandle(h, None, lib))
        return h

    def expt_batch(self, ts: Sequence[float], context: Optional[DeviceContext] = None) -> np.ndarray:
        """e^{Qt} for every t in one kernel launch: ``(len(ts), 4, 4)``; negative t -> 0."""
        ctx = context or default_context()
        return ctx.pmatrix_batch(self._device_handle(ctx), np.asarray(ts, dtype=np.float64))

    def expt(self, t: float) -> np.ndarray:
        """e^{Qt}; the result is also stored as ``self.Qt`` (reference :559-581)."""
        self.Qt = self.expt_batch([t])[0]
        return self.Qt

    def _is_valid(self, transitions: List[float], freqs: List[float], states: int) -> None:
        if len(freqs) != states or not math.isclose(sum(freqs), 1.0, abs_tol=_SUM_TOL):
            raise SubstitutionModelError("Base frequency list either does not sum to 1 or is not of correct length")
        proper_len = ((states - 1) * states) // 2
        if len(transitions) != proper_len:
            raise SubstitutionModelError("Incorrect number of transition rates. Got %d. Expected %d!"
                                         % (len(transitions), proper_len))

    def __getstate__(self) -> Dict[str, Any]:
        d = dict(self.__dict__)
        d["_handle"] = None
        return d


class JC(GTR):
    """Jukes-Cantor: no free parameters."""
    HYPERPARAMS: Tuple[str, ...] = ()

    def __init__(self) -> None:
        super().__init__([0.25] * 4, [1.0] * 6)


class K80(GTR):
    """Kimura 2-parameter: one transition/transversion ratio, uniform frequencies."""
    HYPERPARAMS: Tuple[str, ...] = ("kappa",)

    def __init__(self, kappa: float) -> None:
        self.kappa = _positive("kappa", kappa)
        super().__init__([0.25] * 4, self._rates())

    @property
    def alpha(self) -> float:
        return self.kappa / (self.kappa + 2.0)

    @property
    def beta(self) -> float:
        return 1.0 / (self.kappa + 2.0)

    def _rates(self) -> List[float]:
        k = self.kappa
        return [1.0, k, 1.0, 1.0, k, 1.0]

    def _apply_hyperparams(self, params: Dict[str, Any]) -> None:
        if "kappa" in params:
            self.kappa = _positive("kappa", params["kappa"])
            self.trans = self._rates()


class F81(GTR):
    """Felsenstein 1981: free base frequencies, equal rates."""
    HYPERPARAMS: Tuple[str, ...] = ("base frequencies",)

    def __init__(self, bases: Sequence[float]) -> None:
        super().__init__(bases, [1.0] * 6)

    def _apply_hyperparams(self, params: Dict[str, Any]) -> None:
        if "base frequencies" in params:
            self.freqs = _as_float_list(params["base frequencies"])


class HKY(GTR):
    """Hasegawa-Kishino-Yano: free base frequencies and one ts/tv ratio."""
    HYPERPARAMS: Tuple[str, ...] = ("kappa", "base frequencies")

    def __init__(self, base_freqs: Sequence[float], kappa: float) -> None:
        self.kappa = _positive("kappa", kappa)
        super().__init__(base_freqs, self._rates())

    def _rates(self) -> List[float]:
        k = self.kappa
        return [1.0, k, 1.0, 1.0, k, 1.0]

    def _apply_hyperparams(self, params: Dict[str, Any]) -> None:
        if "kappa" in params:
            self.kappa = _positive("kappa", params["kappa"])
            self.trans = self._rates()
        if "base frequencies" in params:
            self.freqs = _as_float_list(params["base frequencies"])


class TN93(GTR):
    """Tamura-Nei 1993: separate purine / pyrimidine transition ratios."""
    HYPERPARAMS: Tuple[str, ...] = ("kappa_r", "kappa_y", "base frequencies")

    def __init__(self, base_freqs: Sequence[float], kappa_r: float, kappa_y: float) -> None:
        self.kappa_r = _positive("kappa_r", kappa_r)
        self.kappa_y = _positive("kappa_y", kappa_y)
        super().__init__(base_freqs, self._rates())

    def _rates(self) -> List[float]:
        return [1.0, self.kappa_r, 1.0, 1.0, self.kappa_y, 1.0]

    def _apply_hyperparams(self, params: Dict[str, Any]) -> None:
        if "kappa_r" in params:
            self.kappa_r = _positive("kappa_r", params["kappa_r"])
        if "kappa_y" in params:
            self.kappa_y = _positive("kappa_y", params["kappa_y"])
        if "kappa_r" in params or "kappa_y" in params:
            self.trans = self._rates()
        if "base frequencies" in params:
            self.freqs = _as_float_list(params["base frequencies"])


class K81(GTR):
    """Kimura 3-substitution-type model: three rate classes, uniform frequencies."""
    HYPERPARAMS: Tuple[str, ...] = ("alpha", "beta", "gamma")

    def __init__(self, alpha: float, beta: float, gamma: float) -> None:
        self._store_rates(alpha, beta, gamma)
        super().__init__([0.25] * 4, self._rates())

    def _store_rates(self, alpha: float, beta: float, gamma: float) -> None:
        self.alpha = _positive("alpha", alpha)
        self.beta = _positive("beta", beta)
        self.gamma = _positive("gamma", gamma)

    def _rates(self) -> List[float]:
        return [self.beta, self.alpha, self.gamma, self.gamma, self.alpha, self.beta]

    def _apply_hyperparams(self, params: Dict[str, Any]) -> None:
        self._store_rates(params.get("alpha", self.alpha), params.get("beta", self.beta),
                          params.get("gamma", self.gamma))
        self.trans = self._rates()


class SYM(GTR):
    """Symmetric model: six free rates, uniform frequencies."""
    HYPERPARAMS: Tuple[str, ...] = ("transitions",)

    def __init__(self, transitions: Sequence[float]) -> None:
        super().__init__([0.25] * 4, transitions)

    def _apply_hyperparams(self, params: Dict[str, Any]) -> None:
        if "transitions" in params:
            self.trans = _as_float_list(params["transitions"])
