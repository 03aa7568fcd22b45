"""Synthetic inputs for tests and benchmarks: random ultrametric gene trees and
GTR-simulated alignments of the shapes BASELINE.json names.

Host NumPy only (input generation, not part of the timed path).  Own generator:
the reference's simulators (``src/_sim_seq.py``) are out of scope and are not
available on the GPU box; what matters for the benchmark is the shape (taxa x
sites), a GTR process and realistic pattern diversity.  Node heights are spaced
about 0.01 substitutions/site apart (SURVEY section 8d uses a caterpillar species
tree with heights 0.01*k), which reproduces the survey's unique-pattern fractions.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .tree import FlatTree

ASCII_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def gtr_eigen(pi: Sequence[float], rates: Sequence[float]):
    """(evals, U, Uinv) of the normalised GTR rate matrix (host; for simulation only)."""
    pi = np.asarray(pi, dtype=np.float64)
    ac, ag, at_, cg, ct, gt = rates
    r = np.array([[0, ac, ag, at_], [ac, 0, cg, ct], [ag, cg, 0, gt], [at_, ct, gt, 0]], dtype=np.float64)
    Q = r * pi[None, :]
    np.fill_diagonal(Q, -Q.sum(axis=1))
    Q /= -float(pi @ np.diag(Q))
    s = np.sqrt(pi)
    B = (Q * s[:, None]) / s[None, :]
    ev, V = np.linalg.eigh(0.5 * (B + B.T))
    return ev, V / s[:, None], V.T * s[None, :]


def random_trees(M: int, n: int, rng: np.random.Generator, height: Optional[float] = None) -> Tuple[np.ndarray, np.ndarray]:
    """M random binary ultrametric trees with n tips, vectorised over trees.

    Nodes 0..n-1 are tips (labels ``t0``..), n..2n-2 internal in order of creation,
    2n-2 the root, so ``parent[i] > i``.  Returns ``parent (M, 2n-1) int32`` and
    ``blen (M, 2n-1) float64`` (root entry NaN).
    """
    if n < 2:
        raise ValueError("need at least 2 tips")
    if height is None:
        height = 0.01 * (n - 1)
    nn = 2 * n - 1
    rows = np.arange(M)
    active = np.tile(np.arange(n, dtype=np.int32), (M, 1))
    parent = np.full((M, nn), -1, dtype=np.int32)
    node_h = np.zeros((M, nn), dtype=np.float64)
    for k in range(n - 1):
        m = n - k
        i = rng.integers(0, m, size=M)
        j = rng.integers(0, m - 1, size=M)
        j = j + (j >= i)
        a, b = active[rows, i], active[rows, j]
        new = n + k
        parent[rows, a] = new
        parent[rows, b] = new
        node_h[:, new] = height * (k + 1 + rng.uniform(-0.3, 0.3, size=M)) / (n - 1)
        lo, hi = np.minimum(i, j), np.maximum(i, j)
        active[rows, lo] = new
        active[rows, hi] = active[rows, m - 1]
    ph = np.take_along_axis(node_h, np.maximum(parent, 0).astype(np.int64), axis=1)
    blen = ph - node_h
    blen[:, nn - 1] = np.nan
    return parent, blen


def simulate_alignments(parent: np.ndarray, blen: np.ndarray, L: int, pi: Sequence[float], rates: Sequence[float],
                        rng: np.random.Generator) -> np.ndarray:
    """Evolve L sites down each tree under GTR(pi, rates); returns tip states as ASCII
    bytes, shape ``(M, n, L)`` uint8.  Trees must satisfy ``parent[i] > i`` (root last)."""
    M, nn = parent.shape
    n = (nn + 1) // 2
    pi = np.asarray(pi, dtype=np.float64)
    ev, U, Ui = gtr_eigen(pi, rates)
    t = np.nan_to_num(blen, nan=0.0)
    Pm = np.einsum("ik,mnk,kj->mnij", U, np.exp(ev[None, None, :] * t[:, :, None]), Ui)
    cum = np.cumsum(np.clip(Pm, 0.0, None), axis=3)
    cum /= cum[..., 3:4]
    rows = np.arange(M)[:, None]
    states = np.empty((M, nn, L), dtype=np.uint8)
    states[:, nn - 1, :] = rng.choice(4, size=(M, L), p=pi / pi.sum()).astype(np.uint8)
    for i in range(nn - 2, -1, -1):
        pst = states[np.arange(M), parent[:, i]]           # (M, L) parent states
        thr = cum[:, i][rows, pst]                          # (M, L, 4) cumulative rows
        u = rng.random((M, L))
        states[:, i, :] = (u[..., None] > thr[..., :3]).sum(axis=-1).astype(np.uint8)
    return ASCII_ACGT[states[:, :n, :]]


def flat_trees(parent: np.ndarray, blen: np.ndarray) -> List[FlatTree]:
    M, nn = parent.shape
    n = (nn + 1) // 2
    labels = ["t%d" % i for i in range(n)] + [None] * (n - 1)
    return [FlatTree(parent[m], blen[m], labels) for m in range(M)]


def tip_labels(n: int) -> List[str]:
    return ["t%d" % i for i in range(n)]


def to_alignment_dict(mat: np.ndarray) -> dict:
    """(n, L) ASCII byte matrix -> {"t0": "ACG...", ...} as the reference API takes."""
    return {"t%d" % i: mat[i].tobytes().decode("ascii") for i in range(mat.shape[0])}


DEFAULT_PI = (0.1, 0.2, 0.3, 0.4)
DEFAULT_RATES = (1.0, 2.5, 0.7, 1.3, 3.1, 0.9)
