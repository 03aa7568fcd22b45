"""Synthetic inputs for tests and benchmarks: random ultrametric gene trees,
GTR-simulated alignments of the shapes BASELINE.json names, random species networks and gene
trees drawn from the multispecies network coalescent on them.

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


# ---------------------------------------------------------------------------------------------
# Species networks and gene trees under the multispecies network coalescent (inputs of K-m)
# ---------------------------------------------------------------------------------------------
def random_species_network(n_species: int, n_retic: int, rng: np.random.Generator, height: float = 0.05):
    """A random ultrametric species network: a coalescent-shaped tree over ``s0..s{n-1}`` plus
    ``n_retic`` reticulations, each a hybrid node placed on a recipient edge with a second parent edge
    coming from a node placed on a donor edge at or above the same height.  Returns a
    :class:`~phynetpy_b200._msnc_density.SpeciesNetwork`."""
    from ._msnc_density import SpeciesNetwork
    label: List[Optional[str]] = ["s%d" % i for i in range(n_species)]
    h: List[float] = [0.0] * n_species
    edges: List[List] = []     # [src, dst, gamma | None]
    live = list(range(n_species))
    t = 0.0
    while len(live) > 1:
        t += float(rng.exponential(height / n_species))
        i, j = sorted(rng.choice(len(live), size=2, replace=False).tolist(), reverse=True)
        a, b = live.pop(i), live.pop(j)
        label.append(None)
        h.append(t)
        v = len(h) - 1
        edges += [[v, a, None], [v, b, None]]
        live.append(v)

    def split(e: int, m: int, gamma_into_m):
        p, c, g = edges[e]
        edges[e] = [p, m, gamma_into_m]
        edges.append([m, c, g])   # a gamma on p -> c stays on the half that still points into c

    for _ in range(n_retic):
        for _try in range(1000):
            r, d = int(rng.integers(len(edges))), int(rng.integers(len(edges)))
            if r == d:
                continue
            (pr, cr, _g), (pd, cd, _g2) = edges[r], edges[d]
            lo_r, hi_r = h[cr], h[pr]
            if hi_r - lo_r < 1e-9:
                continue
            th = float(rng.uniform(lo_r + 0.1 * (hi_r - lo_r), hi_r - 0.1 * (hi_r - lo_r)))
            lo_d, hi_d = max(h[cd], th), h[pd]
            if hi_d - lo_d < 1e-9:
                continue
            tx = float(rng.uniform(lo_d + 0.05 * (hi_d - lo_d), hi_d - 0.05 * (hi_d - lo_d)))
            g = float(rng.uniform(0.1, 0.9))
            label.extend([None, None])
            h.extend([th, tx])
            hyb, don = len(h) - 2, len(h) - 1
            split(r, hyb, 1.0 - g)
            split(d, don, None)
            edges.append([don, hyb, g])
            break
        else:
            raise RuntimeError("could not place a reticulation")
    return SpeciesNetwork([e[0] for e in edges], [e[1] for e in edges], [e[2] for e in edges], h, label)


def simulate_gene_trees_msnc(net, alleles_per_species: int, theta: float, rng: np.random.Generator, M: int):
    """M gene trees drawn from the multispecies network coalescent on ``net`` (the process the timed
    MSNC density scores): lineages enter at the leaves, coalesce pairwise at rate u(u-1)/theta inside a
    branch, choose a parent edge with probability gamma at a reticulation, and finish above the root.

    Returns ``(trees, species_of)``: :class:`FlatTree` objects with leaves first and the root last
    (``parent[i] > i``, so :func:`simulate_alignments` accepts their stacked arrays) and the
    allele -> species map.  Allele labels are ``t0, t1, ...`` in species order."""
    n_nodes = net.n_nodes
    down = [[] for _ in range(n_nodes)]
    up = [[] for _ in range(n_nodes)]
    for e, (s, d) in enumerate(zip(net.edge_src.tolist(), net.edge_dst.tolist())):
        down[s].append(e)
        up[d].append(e)
    gam = net.edge_gamma.tolist()
    order = sorted(range(n_nodes), key=lambda v: (net.height[v], len(down[v]) > 0))
    n_leaves = len(net.species) * alleles_per_species
    labels = ["t%d" % i for i in range(n_leaves)] + [None] * (n_leaves - 1)
    species_of = {"t%d" % (s * alleles_per_species + k): net.species[s]
                  for s in range(len(net.species)) for k in range(alleles_per_species)}
    trees = []
    for _ in range(M):
        parent = [-1] * (2 * n_leaves - 1)
        node_h = [0.0] * (2 * n_leaves - 1)
        nxt = n_leaves
        at_edge = {}     # edge -> lineages (gene-tree node ids) waiting at its bottom

        def coalesce(lin, lo, hi):
            nonlocal nxt
            t = lo
            lin = list(lin)
            while len(lin) > 1:
                u = len(lin)
                t += float(rng.exponential(theta / (u * (u - 1))))
                if hi is not None and t >= hi:
                    break
                i, j = sorted(rng.choice(u, size=2, replace=False).tolist(), reverse=True)
                a, b = lin.pop(i), lin.pop(j)
                parent[a] = parent[b] = nxt
                node_h[nxt] = t
                lin.append(nxt)
                nxt += 1
            return lin

        for v in order:
            if not down[v]:
                s = int(net.node_species[v])
                arriving = list(range(s * alleles_per_species, (s + 1) * alleles_per_species))
            else:
                arriving = []
                for e in down[v]:
                    arriving += coalesce(at_edge.pop(e, []), float(net.height[net.edge_dst[e]]), float(net.height[v]))
            if not up[v]:
                coalesce(arriving, float(net.height[v]), None)
            elif len(up[v]) == 2:
                e1, e2 = up[v]
                g1, g2 = gam[e1], gam[e2]
                if g1 != g1 and g2 != g2:
                    g1 = 0.5
                elif g1 != g1:
                    g1 = 1.0 - g2
                for x in arriving:
                    at_edge.setdefault(e1 if rng.random() < g1 else e2, []).append(x)
            else:
                at_edge.setdefault(up[v][0], []).extend(arriving)
        blen = [float("nan") if parent[i] < 0 else node_h[parent[i]] - node_h[i] for i in range(2 * n_leaves - 1)]
        trees.append(FlatTree(parent, blen, labels))
    return trees, species_of


# ---------------------------------------------------------------------------------------------
# Counter-based sequence simulation: on the device (K-s, pnb_simulate_alignments) and its NumPy twin
# ---------------------------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic), as ``sim_mix`` in csrc/pnb_simulate.cu."""
    with np.errstate(over="ignore"):
        x = x ^ (x >> np.uint64(30))
        x = x * np.uint64(0xBF58476D1CE4E5B9)
        x = x ^ (x >> np.uint64(27))
        x = x * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def counter_uniform(seed: int, stream: int, sites: np.ndarray, node: int) -> np.ndarray:
    """U[0, 1) draws keyed by (seed, stream, site, node): the generator of K-s, bit for bit."""
    with np.errstate(over="ignore"):
        s0 = _mix64(np.array([seed], dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15) * np.uint64(stream + 1))
        k = sites.astype(np.uint64) * np.uint64(0xD1B54A32D192ED03) + np.uint64(node) * np.uint64(0x8CB92BA72F3D8DD7) \
            + np.uint64(0x2545F4914F6CDD1D)
        x = _mix64(s0 ^ k)
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def edge_p_matrices(blen: np.ndarray, pi: Sequence[float], rates: Sequence[float]) -> np.ndarray:
    """Row-major P(t) of the edge above every node, ``(..., 16)`` float64, on the host (NaN length -> 0): the
    thresholds both simulators compare against."""
    ev, U, Ui = gtr_eigen(pi, rates)
    t = np.nan_to_num(np.asarray(blen, dtype=np.float64), nan=0.0)
    Pm = np.einsum("ik,...k,kj->...ij", U, np.exp(ev * t[..., None]), Ui)
    return np.ascontiguousarray(Pm.reshape(t.shape + (16,)))


def _cum_rows(P16: np.ndarray) -> np.ndarray:
    """Sequential cumulative sums of the clipped rows (the order ``cum_rows`` uses on the device side)."""
    P = np.clip(P16.reshape(-1, 4, 4), 0.0, None)
    c = np.empty_like(P)
    s = np.zeros(P.shape[:2])
    for j in range(4):
        s = s + P[:, :, j]
        c[:, :, j] = s
    return c


def simulate_alignments_counter(trees: Sequence[FlatTree], L: int, pi: Sequence[float], rates: Sequence[float], seed: int,
                                stream_ids: Sequence[int]) -> List[np.ndarray]:
    """NumPy twin of K-s: the alignment of every tree as an ``(n_tips, L)`` ASCII matrix (leaves in node order)."""
    pi = np.asarray(pi, dtype=np.float64)
    c = np.cumsum(pi)          # sequential: c0, c1, c2 as on the device
    sites = np.arange(L)
    out = []
    for t, sid in zip(trees, stream_ids):
        nn = t.n_nodes
        cum = _cum_rows(edge_p_matrices(t.blen, pi, rates))
        kids = [[] for _ in range(nn)]
        root = -1
        for v, p in enumerate(t.parent.tolist()):
            if p < 0:
                root = v
            else:
                kids[p].append(v)
        st = np.zeros((nn, L), dtype=np.uint8)
        stack = [root]
        while stack:
            v = stack.pop()
            u = counter_uniform(seed, int(sid), sites, v)
            if v == root:
                st[v] = (u > c[0]).astype(np.uint8) + (u > c[1]) + (u > c[2])
            else:
                row = cum[v][st[t.parent[v]]]          # (L, 4)
                st[v] = (u > row[:, 0]).astype(np.uint8) + (u > row[:, 1]) + (u > row[:, 2])
            stack.extend(kids[v])
        tips = [v for v in range(nn) if not kids[v]]
        out.append(ASCII_ACGT[st[tips]])
    return out


def simulate_alignments_device(trees: Sequence[FlatTree], L: int, pi: Sequence[float], rates: Sequence[float], seed: int,
                               stream_ids: Sequence[int], context=None) -> List[np.ndarray]:
    """K-s through the C ABI: every tree's alignment in one launch (one thread per tree and site)."""
    import ctypes as C
    from . import _lib
    from .device import default_context
    ctx = context or default_context()
    M = len(trees)
    if M == 0:
        return []
    sizes = np.fromiter((t.n_nodes for t in trees), dtype=np.int64, count=M)
    node_off = np.zeros(M + 1, dtype=np.int64)
    np.cumsum(sizes, out=node_off[1:])
    parent = np.ascontiguousarray(np.concatenate([t.parent for t in trees]), dtype=np.int32)
    blen = np.concatenate([t.blen for t in trees])
    P_edge = edge_p_matrices(blen, pi, rates)
    n_tips = np.fromiter((int((t.n_children() == 0).sum()) for t in trees), dtype=np.int64, count=M)
    out = np.empty(int((n_tips * L).sum()), dtype=np.uint8)
    out_off = np.zeros(M + 1, dtype=np.int64)
    sid = np.ascontiguousarray(stream_ids, dtype=np.int64)
    pi_a = np.ascontiguousarray(pi, dtype=np.float64)
    _lib.check(ctx._lib.pnb_simulate_alignments(ctx.handle, M, _lib.ptr(node_off), _lib.ptr(parent), _lib.ptr(P_edge), _lib.ptr(pi_a),
                                                int(L), C.c_uint64(int(seed)), _lib.ptr(sid), _lib.ptr(out), _lib.ptr(out_off)))
    return [out[out_off[j]:out_off[j + 1]].reshape(int(n_tips[j]), L) for j in range(M)]
