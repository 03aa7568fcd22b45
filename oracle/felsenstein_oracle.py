"""CPU oracle for the sequence-likelihood hot path -- TEST INFRASTRUCTURE ONLY.

A plain NumPy restatement of the reference algorithm (PhyNetPy v0.6.0,
``src/_seq_likelihood.py``), function by function, used as the checker in
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py``.  Nothing in the product package imports it and
the product never falls back to it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` ran the unmodified
reference in the build container and froze its outputs (pattern counts, weights,
P(t) matrices, log-likelihoods) into ``tests/golden/*.json``;
``tests/test_oracle_golden.py`` checks this file against every one of them, and
against the reference's own known-answer test (``tests/test_mcmc_seq.py:107-121``:
pruning == per-column ``scipy.linalg.expm`` pruning).

Every function cites the reference lines it follows (paths relative to the
reference checkout).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# --- encoding: src/_seq_likelihood.py:104-117 ------------------------------
DNA_STATES = "ACGT"
_STATE_INDEX = {"A": 0, "C": 1, "G": 2, "T": 3}
_IUPAC = {
    "A": "A", "C": "C", "G": "G", "T": "T", "U": "T",
    "R": "AG", "Y": "CT", "S": "CG", "W": "AT", "K": "GT", "M": "AC",
    "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG",
    "N": "ACGT", "?": "ACGT", "-": "ACGT", "X": "ACGT", ".": "ACGT",
    "O": "ACGT",
}


def tip_partials_for_char(ch: str) -> np.ndarray:
    """src/_seq_likelihood.py:120-140 -- unknown characters are all ones."""
    out = np.zeros(4, dtype=np.float64)
    allowed = _IUPAC.get(str(ch).upper())
    if allowed is None:
        out[:] = 1.0
        return out
    for nt in allowed:
        out[_STATE_INDEX[nt]] = 1.0
    return out


def tip_code_for_char(ch: str) -> int:
    """4-bit mask view of :func:`tip_partials_for_char` (A=1, C=2, G=4, T=8)."""
    v = tip_partials_for_char(ch)
    return int(v[0]) | int(v[1]) << 1 | int(v[2]) << 2 | int(v[3]) << 3


# --- substitution models: src/_seq_likelihood.py:162-304 -------------------
class OracleModel:
    """Q, eigen-system and P(t) exactly as ``SubstitutionModel`` builds them."""

    def __init__(self, pi: Sequence[float], rates: Sequence[float], jc_closed_form: bool = False):
        # :171-177
        self.pi = np.asarray(pi, dtype=np.float64)
        if self.pi.shape != (4,):
            raise ValueError("Base frequencies must have length 4 (A,C,G,T).")
        if not math.isclose(float(self.pi.sum()), 1.0, abs_tol=1e-9):
            raise ValueError("Base frequencies must sum to 1.")
        self._rates = np.asarray(rates, dtype=np.float64)
        self.jc_closed_form = jc_closed_form
        # :189-200
        ac, ag, at_, cg, ct, gt = self._rates
        r = np.array([[0.0, ac, ag, at_], [ac, 0.0, cg, ct], [ag, cg, 0.0, gt], [at_, ct, gt, 0.0]],
                     dtype=np.float64)
        Q = r * self.pi[np.newaxis, :]
        # :204-214
        np.fill_diagonal(Q, 0.0)
        np.fill_diagonal(Q, -Q.sum(axis=1))
        mean_rate = -float(np.dot(self.pi, np.diag(Q)))
        if mean_rate <= 0.0:
            raise ValueError("Degenerate substitution model (non-positive rate).")
        Q /= mean_rate
        self.Q = Q
        # :219-227
        sqrt_pi = np.sqrt(self.pi)
        inv_sqrt_pi = 1.0 / sqrt_pi
        B = (Q * sqrt_pi[:, None]) * inv_sqrt_pi[None, :]
        B = 0.5 * (B + B.T)
        evals, evecs = np.linalg.eigh(B)
        self._evals = evals
        self._U = inv_sqrt_pi[:, None] * evecs
        self._Uinv = evecs.T * sqrt_pi[None, :]

    def p_matrix(self, t: float) -> np.ndarray:
        if self.jc_closed_form:
            # JC69.p_matrix, :259-268
            if t < 0.0:
                t = 0.0
            e = math.exp(-4.0 * t / 3.0)
            P = np.full((4, 4), 0.25 - 0.25 * e, dtype=np.float64)
            np.fill_diagonal(P, 0.25 + 0.75 * e)
            return P
        # SubstitutionModel.p_matrix, :241-244
        if t < 0.0:
            t = 0.0
        expd = np.exp(self._evals * t)
        return (self._U * expd[None, :]) @ self._Uinv


def jc69() -> OracleModel:
    return OracleModel([0.25] * 4, [1.0] * 6, jc_closed_form=True)  # :255-257


def hky85(kappa: float, pi: Sequence[float]) -> OracleModel:
    k = float(kappa)
    return OracleModel(pi, [1.0, k, 1.0, 1.0, k, 1.0])  # :285-286


def gtr(pi: Sequence[float], rates: Sequence[float]) -> OracleModel:
    return OracleModel(pi, rates)  # :297-304


# --- trees as flat arrays ---------------------------------------------------
class OracleTree:
    """Rooted tree: ``parent[i]`` (-1 root), ``blen[i]`` (None allowed), ``label[i]``.

    ``children[i]`` keeps insertion order; the reference iterates a Python set
    (``graph_core_cy.pyx:128-135``), i.e. an arbitrary order, which matters only
    in the last ulp for out-degree >= 3.
    """

    def __init__(self, parent: Sequence[int], blen: Sequence[Optional[float]], label: Sequence[Optional[str]]):
        self.parent = list(parent)
        self.blen = list(blen)
        self.label = list(label)
        n = len(self.parent)
        self.children: List[List[int]] = [[] for _ in range(n)]
        roots = []
        for i, p in enumerate(self.parent):
            if p < 0:
                roots.append(i)
            else:
                self.children[p].append(i)
        if len(roots) != 1:
            raise ValueError("tree must have exactly one root")
        self.root = roots[0]

    @staticmethod
    def from_newick(s: str) -> "OracleTree":
        """Minimal Newick reader (labels, branch lengths, nesting); own code, not the reference's."""
        s = s.strip()
        if s.endswith(";"):
            s = s[:-1]
        parent: List[int] = []
        blen: List[Optional[float]] = []
        label: List[Optional[str]] = []

        def new_node(par: int) -> int:
            parent.append(par)
            blen.append(None)
            label.append(None)
            return len(parent) - 1

        pos = 0
        cur = new_node(-1)
        stack: List[int] = []
        # recursive-descent without recursion: '(' opens a child of cur, ',' a sibling, ')' closes
        while pos < len(s):
            ch = s[pos]
            if ch == "(":
                stack.append(cur)
                cur = new_node(cur)
                pos += 1
            elif ch == ",":
                cur = new_node(stack[-1])
                pos += 1
            elif ch == ")":
                cur = stack.pop()
                pos += 1
            elif ch == ":":
                pos += 1
                start = pos
                while pos < len(s) and s[pos] not in ",()":
                    pos += 1
                blen[cur] = float(s[start:pos])
            elif ch.isspace():
                pos += 1
            else:
                start = pos
                while pos < len(s) and s[pos] not in ":,()":
                    pos += 1
                label[cur] = s[start:pos].strip()
        return OracleTree(parent, blen, label)


# --- pattern compression: src/_seq_likelihood.py:336-367 -------------------
class OracleLocus:
    def __init__(self, alignment: Dict[str, str]):
        if not alignment:
            raise ValueError("Empty alignment passed to FelsensteinCalculator.")  # :336-337
        self.taxa = list(alignment.keys())  # :338
        seqs = [alignment[t] for t in self.taxa]
        length = len(seqs[0])
        if any(len(s) != length for s in seqs):
            raise ValueError("All sequences in a locus must be equally long.")  # :341-342
        pattern_index: Dict[tuple, int] = {}
        weights: List[int] = []
        columns: List[tuple] = []
        pattern_of_col: List[int] = []
        for col in range(length):  # :348-356
            key = tuple(seqs[r][col].upper() for r in range(len(self.taxa)))
            idx = pattern_index.get(key)
            if idx is None:
                idx = len(columns)
                pattern_index[key] = idx
                columns.append(key)
                weights.append(1)
            else:
                weights[idx] += 1
            pattern_of_col.append(idx)
        self.columns = columns
        self.n_patterns = len(columns)  # :358
        self.int_weights = np.asarray(weights, dtype=np.int64)
        self.weights = np.asarray(weights, dtype=np.float64)  # :359
        self.pattern_of_col = np.asarray(pattern_of_col, dtype=np.int64)
        self.tip_partials: Dict[str, np.ndarray] = {}
        for r, taxon in enumerate(self.taxa):  # :362-367
            mat = np.empty((self.n_patterns, 4), dtype=np.float64)
            for p, key in enumerate(columns):
                mat[p] = tip_partials_for_char(key[r])
            self.tip_partials[taxon] = mat

    def codes(self) -> np.ndarray:
        """(n_taxa, n_patterns) uint8 4-bit masks equivalent to ``tip_partials``."""
        out = np.zeros((len(self.taxa), self.n_patterns), dtype=np.uint8)
        for r in range(len(self.taxa)):
            for p, key in enumerate(self.columns):
                out[r, p] = tip_code_for_char(key[r])
        return out


def compress_fast(seq_matrix: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Vectorised first-occurrence compression of an (n, L) uint8 matrix.

    Same numbering as :class:`OracleLocus` (checked in the tests) but usable at
    10^6 columns: returns (first_col, weights, pattern_of_col).
    """
    n, L = seq_matrix.shape
    colsT = np.ascontiguousarray(seq_matrix.T)
    view = colsT.view(np.dtype((np.void, n))).ravel()
    _, first, inv, cnt = np.unique(view, return_index=True, return_inverse=True, return_counts=True)
    order = np.argsort(first, kind="stable")  # unique groups ordered by first occurrence
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return first[order].astype(np.int64), cnt[order].astype(np.int64), rank[inv].astype(np.int64)


# --- pruning: src/_seq_likelihood.py:369-426 --------------------------------
def postorder_partials(locus_tip: Dict[str, np.ndarray], n_patterns: int, tree: OracleTree, node: int,
                       model: OracleModel, site_rate: float, sink: Optional[dict] = None) -> np.ndarray:
    """``_postorder_partials`` (:399-426); ``sink`` collects every node's partial."""
    children = tree.children[node]
    if not children:
        tip = locus_tip.get(tree.label[node])
        if tip is None:
            out = np.ones((n_patterns, 4), dtype=np.float64)  # :409-412
        else:
            out = tip
        if sink is not None:
            sink[node] = out
        return out
    result = np.ones((n_patterns, 4), dtype=np.float64)  # :415
    for child in children:
        t = tree.blen[child]
        t = 0.0 if t is None else float(t) * site_rate  # :419-420
        child_partials = postorder_partials(locus_tip, n_patterns, tree, child, model, site_rate, sink)
        P = model.p_matrix(t)  # :422
        result *= child_partials @ P.T  # :425
    if sink is not None:
        sink[node] = result
    return result


def log_likelihood(locus: OracleLocus, tree: OracleTree, model: OracleModel, site_rate: float = 1.0,
                   sink: Optional[dict] = None) -> float:
    """``FelsensteinCalculator.log_likelihood`` (:391-397)."""
    partials = postorder_partials(locus.tip_partials, locus.n_patterns, tree, tree.root, model, site_rate, sink)
    site_like = partials @ model.pi
    site_like = np.maximum(site_like, 1e-300)
    return float(np.dot(np.log(site_like), locus.weights))


def log_likelihood_codes(codes: np.ndarray, weights: np.ndarray, tip_row: Sequence[int], tree: OracleTree,
                         model: OracleModel, site_rate: float = 1.0) -> float:
    """Same computation from (n_rows, P) 4-bit tip codes (large synthetic loci).

    ``tip_row[i]`` is the code row of leaf ``i`` (-1: taxon absent -> all ones).
    """
    n_patterns = codes.shape[1]
    bits = np.array([[(c >> j) & 1 for j in range(4)] for c in range(16)], dtype=np.float64)
    tips = {}
    label = list(tree.label)
    for i in range(len(tree.parent)):
        if not tree.children[i]:
            r = tip_row[i]
            name = "__tip%d" % i
            label[i] = name
            if r >= 0:
                tips[name] = bits[codes[r]]
    t2 = OracleTree(tree.parent, tree.blen, label)
    partials = postorder_partials(tips, n_patterns, t2, t2.root, model, site_rate)
    site_like = np.maximum(partials @ model.pi, 1e-300)
    return float(np.dot(np.log(site_like), np.asarray(weights, dtype=np.float64)))


class OracleCodesLocus:
    """Tip tables prebuilt from (n_rows, P) codes, as the reference's constructor leaves them
    (:361-367), so that timing an evaluation excludes construction."""

    def __init__(self, codes: np.ndarray, weights: np.ndarray):
        bits = np.array([[(c >> j) & 1 for j in range(4)] for c in range(16)], dtype=np.float64)
        self.n_patterns = int(codes.shape[1])
        self.rows = [np.ascontiguousarray(bits[codes[r]]) for r in range(codes.shape[0])]
        self.weights = np.asarray(weights, dtype=np.float64)

    def log_likelihood(self, tip_row: Sequence[int], tree: OracleTree, model: OracleModel, site_rate: float = 1.0) -> float:
        tips = {}
        label = list(tree.label)
        for i in range(len(tree.parent)):
            if not tree.children[i]:
                label[i] = "__tip%d" % i
                if tip_row[i] >= 0:
                    tips[label[i]] = self.rows[tip_row[i]]
        t2 = OracleTree(tree.parent, tree.blen, label)
        partials = postorder_partials(tips, self.n_patterns, t2, t2.root, model, site_rate)
        site_like = np.maximum(partials @ model.pi, 1e-300)
        return float(np.dot(np.log(site_like), self.weights))


def log_likelihood_rescaled(locus: OracleLocus, tree: OracleTree, model: OracleModel, site_rate: float = 1.0) -> float:
    """NOT in the reference: the same pruning with per-node renormalisation (each node divided by its
    largest state, the logs accumulated), i.e. what the likelihood is where the reference underflows.
    Independent check for the library's optional PNB_EVAL_RESCALE mode."""
    n_pat = locus.n_patterns
    logscale = np.zeros(n_pat)

    def rec(node: int) -> np.ndarray:
        nonlocal logscale
        kids = tree.children[node]
        if not kids:
            tip = locus.tip_partials.get(tree.label[node])
            return np.ones((n_pat, 4)) if tip is None else tip
        res = np.ones((n_pat, 4))
        for c in kids:
            t = tree.blen[c]
            t = 0.0 if t is None else float(t) * site_rate
            res = res * (rec(c) @ model.p_matrix(t).T)
        m = res.max(axis=1)
        ok = m > 0
        res[ok] /= m[ok, None]
        logscale = logscale + np.where(ok, np.log(np.where(ok, m, 1.0)), 0.0)
        return res

    import sys
    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 10 * len(tree.parent) + 100))
    try:
        root = rec(tree.root)
    finally:
        sys.setrecursionlimit(old)
    site = root @ model.pi
    with np.errstate(divide="ignore"):
        ls = np.where(site > 0, np.log(np.where(site > 0, site, 1.0)) + logscale, math.log(1e-300))
    return float(np.dot(ls, locus.weights))


# --- engine loop: src/_mcmc_seq.py:746-764 (Felsenstein factor only) -------
def engine_log_likelihood(loci: Sequence[OracleLocus], trees: Sequence[OracleTree], model: OracleModel) -> Tuple[float, List[float]]:
    per_locus = [log_likelihood(l, t, model) for l, t in zip(loci, trees)]
    total = 0.0
    for v in per_locus:
        total += v
    if not math.isfinite(total):
        return float("-inf"), per_locus
    return total, per_locus
