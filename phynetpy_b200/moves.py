"""Gene-tree proposals on flat arrays ("flat-array gene-tree mirror updated by
the moves").

The reference's gene-tree operators clone the ``Network`` of the chosen locus, recompute every node
height, mutate the clone and re-derive every edge length (``op_gene_node_height`` /
``op_gene_tree_nni``, ``src/_mcmc_seq.py:1020-1113``; ``_clone_net`` alone is ~0.6 ms, about 30% of the
chain's time).  Here a gene tree is a :class:`~phynetpy_b200.tree.FlatTree` plus its
node-height vector, and a proposal is a handful of NumPy operations that produce the next
``(FlatTree, heights)`` pair -- the old pair stays untouched, which is the copy-and-swap convention the
likelihood engine relies on ("a fresh object means a genuine change", ``src/_mcmc_seq.py:682-694``).

Same proposal distributions as the reference:

* :func:`slide_height_move` (``_slide_height_move``, ``:956-990``): candidates are the non-root internal
  nodes plus the root; a non-root node slides uniformly inside (max child height, parent height), log
  Hastings ratio 0; the root is scaled by ``f = exp(w (u - 1/2))``, log Hastings ratio ``log f``;
* :func:`nni_move` (``op_gene_tree_nni``, ``:1053-1113``): for every non-root internal node ``v`` one of its
  children ``c`` is drawn, then one ``(p, v, s, c)`` uniformly; ``c`` and ``v``'s sibling ``s`` trade places
  when ``h[v] > h[s]``; heights are kept, so the move is symmetric;
* edge lengths always follow ``max(h[parent] - h[child], 0)`` (``_sync_lengths``, ``:418-434``).

Random numbers are drawn in the reference's order (node choice, then the uniform), from the caller's
``numpy.random.Generator``.  Which NODE a given integer selects differs: the reference enumerates
``net.V()``, a hash-ordered set whose order changes from clone to clone, here nodes are enumerated by
index -- the distributions are identical, the streams are not comparable.  The deterministic halves
(:func:`slide_node`, :func:`nni_swap`) are what the golden vectors pin.

The parameter moves of the species network have flat twins too -- :func:`net_slide_height_move`
(``op_net_node_height``, ``:993-1017``; the same ``_slide_height_move`` on the network, a reticulation being bounded
by BOTH parents), :func:`gamma_move` (``op_change_gamma``, ``:909-953``) and :func:`theta_move`
(``op_change_theta``, ``:886-906``) -- each returning a fresh :class:`~phynetpy_b200._msnc_density.SpeciesNetwork`
(or theta) so that the engine's identity-keyed network cache sees a genuine change.  The dimension-changing
moves (add / delete / relocate reticulation) stay with the sampler; their gene-tree half is ``coupled.py``.
"""
from __future__ import annotations

import math
from typing import Any, List, Optional, Tuple

import numpy as np

from .tree import FlatTree


def lengths_from_heights(parent: np.ndarray, h: np.ndarray) -> np.ndarray:
    """``_sync_lengths``: every edge gets ``max(h[parent] - h[child], 0)``; the root's entry is NaN."""
    out = np.full(parent.shape[0], np.nan, dtype=np.float64)
    nz = parent >= 0
    out[nz] = np.maximum(h[parent[nz]] - h[nz], 0.0)
    return out


def _rebuilt(ft: FlatTree, parent: np.ndarray, h: np.ndarray, same_topology: bool) -> FlatTree:
    t = FlatTree.__new__(FlatTree)
    t.parent = parent
    t.blen = lengths_from_heights(parent, h)
    t.label = ft.label
    t._n_children = ft._n_children if same_topology else None
    t._tip_rows = ft._tip_rows          # leaves stay leaves under both moves
    return t


def slide_candidates(ft: FlatTree) -> np.ndarray:
    """Non-root internal nodes in index order, then the root (``_internal_nodes(net) + [root]``)."""
    nk = ft.n_children()
    internal = np.nonzero((nk > 0) & (ft.parent >= 0))[0]
    root = np.nonzero(ft.parent < 0)[0][:1]
    return np.concatenate([internal, root]).astype(np.int64)


def slide_node(ft: FlatTree, h: np.ndarray, node: int, u: float,
               root_scale_window: float = 0.3) -> Optional[Tuple[np.ndarray, float]]:
    """The deterministic half of the slide: new height vector and log Hastings ratio for ``node`` given
    the uniform ``u``; ``None`` when the move is impossible (``:971-990``)."""
    kids = np.nonzero(ft.parent == node)[0]
    lower = float(h[kids].max()) if kids.shape[0] else 0.0
    p = int(ft.parent[node])
    h2 = h.copy()
    if p < 0:
        f = math.exp(root_scale_window * (u - 0.5))
        new_h = float(h[node]) * f
        if new_h <= lower:
            return None
        h2[node] = new_h
        return h2, math.log(f)
    upper = float(h[p])
    if upper <= lower:
        return None
    h2[node] = lower + (upper - lower) * u
    return h2, 0.0


def slide_height_move(ft: FlatTree, h: np.ndarray, rng: np.random.Generator,
                      root_scale_window: float = 0.3) -> Optional[Tuple[FlatTree, np.ndarray, float]]:
    """``op_gene_node_height`` without the graph: returns the proposed ``(tree, heights, log HR)``."""
    cands = slide_candidates(ft)
    if cands.shape[0] == 0:
        return None
    node = int(cands[int(rng.integers(cands.shape[0]))])
    out = slide_node(ft, h, node, float(rng.random()), root_scale_window)
    if out is None:
        return None
    h2, loghr = out
    return _rebuilt(ft, ft.parent, h2, True), h2, loghr


def nni_eligible(ft: FlatTree) -> List[Tuple[int, int, int]]:
    """``(p, v, s)`` for every non-root node ``v`` with at least two children whose parent ``p`` has
    another child; ``s`` is the first such sibling (``:1071-1088``)."""
    nk = ft.n_children()
    out = []
    for v in np.nonzero((nk >= 2) & (ft.parent >= 0))[0].tolist():
        p = int(ft.parent[v])
        sibs = np.nonzero(ft.parent == p)[0]
        sibs = sibs[sibs != v]
        if sibs.shape[0]:
            out.append((p, v, int(sibs[0])))
    return out


def nni_swap(ft: FlatTree, h: np.ndarray, v: int, c: int) -> Optional[FlatTree]:
    """The deterministic half of the NNI: child ``c`` of ``v`` and the sibling ``s`` of ``v`` trade places
    (``s`` goes under ``v``, ``c`` under ``v``'s parent); ``None`` unless ``h[v] > h[s]`` (``:1094-1104``)."""
    p = int(ft.parent[v])
    if p < 0 or int(ft.parent[c]) != v:
        raise ValueError("nni_swap: v must be a non-root node and c one of its children")
    sibs = np.nonzero(ft.parent == p)[0]
    sibs = sibs[sibs != v]
    if sibs.shape[0] == 0:
        raise ValueError("nni_swap: v has no sibling")
    s = int(sibs[0])
    if h[v] <= h[s]:
        return None
    parent = ft.parent.copy()
    parent[s], parent[c] = v, p
    return _rebuilt(ft, parent, h, False)


def nni_move(ft: FlatTree, h: np.ndarray, rng: np.random.Generator) -> Optional[Tuple[FlatTree, np.ndarray]]:
    """``op_gene_tree_nni`` without the graph: one child is drawn for EVERY eligible node (as the reference
    does while it builds its list), then one eligible tuple; returns the proposed ``(tree, heights)``."""
    picks = []
    for p, v, s in nni_eligible(ft):
        kids = np.nonzero(ft.parent == v)[0]
        picks.append((v, int(kids[int(rng.integers(kids.shape[0]))])))
    if not picks:
        return None
    v, c = picks[int(rng.integers(len(picks)))]
    t = nni_swap(ft, h, v, c)
    if t is None:
        return None
    return t, h


# ---------------------------------------------------------------------------------------------------
# species-network parameter moves
# ---------------------------------------------------------------------------------------------------
def theta_move(theta: float, rng: np.random.Generator, window: float = 0.4) -> Tuple[float, float]:
    """``op_change_theta``: ``theta * f`` with ``f = exp(window (u - 1/2))``; returns ``(theta', log f)``."""
    f = math.exp(window * (float(rng.random()) - 0.5))
    return theta * f, math.log(f)


def net_slide_candidates(net: Any) -> np.ndarray:
    """Non-root internal network nodes in index order, then the root."""
    n = net.n_nodes
    out_deg = np.bincount(net.edge_src, minlength=n)
    in_deg = np.bincount(net.edge_dst, minlength=n)
    internal = np.nonzero((out_deg > 0) & (in_deg > 0))[0]
    root = np.nonzero(in_deg == 0)[0][:1]
    return np.concatenate([internal, root]).astype(np.int64)


def net_slide_node(net: Any, node: int, u: float, root_scale_window: float = 0.3) -> Optional[Tuple[Any, float]]:
    """The deterministic half of ``op_net_node_height``: ``node`` slides uniformly between its tallest child and its
    LOWEST parent (a reticulation has two), the root is scaled; ``None`` when the move is impossible."""
    h = net.height
    kids = net.edge_dst[net.edge_src == node]
    parents = net.edge_src[net.edge_dst == node]
    lower = float(h[kids].max()) if kids.shape[0] else 0.0
    h2 = h.copy()
    if parents.shape[0] == 0:
        f = math.exp(root_scale_window * (u - 0.5))
        new_h = float(h[node]) * f
        if new_h <= lower:
            return None
        h2[node] = new_h
        return net.replaced(height=h2), math.log(f)
    upper = float(h[parents].min())
    if upper <= lower:
        return None
    h2[node] = lower + (upper - lower) * u
    return net.replaced(height=h2), 0.0


def net_slide_height_move(net: Any, rng: np.random.Generator, root_scale_window: float = 0.3) -> Optional[Tuple[Any, float]]:
    """``op_net_node_height`` without the graph: returns the proposed ``(network, log HR)``."""
    cands = net_slide_candidates(net)
    if cands.shape[0] == 0:
        return None
    node = int(cands[int(rng.integers(cands.shape[0]))])
    return net_slide_node(net, node, float(rng.random()), root_scale_window)


def reticulations(net: Any) -> np.ndarray:
    return np.nonzero(np.bincount(net.edge_dst, minlength=net.n_nodes) >= 2)[0].astype(np.int64)


def gamma_move_at(net: Any, retic: int, u: float, window: float = 0.2, first_edge: Optional[int] = None) -> Any:
    """The deterministic half of ``op_change_gamma``: the inheritance probability of one parent edge of ``retic``
    (``first_edge``, by default the first in edge order) takes a step ``window (u - 1/2)``, reflected into (0, 1);
    the other parent edge gets the complement."""
    edges = np.nonzero(net.edge_dst == retic)[0]
    if edges.shape[0] < 2:
        raise ValueError("gamma_move_at: node %d is not a reticulation" % retic)
    e0 = int(edges[0]) if first_edge is None else int(first_edge)
    e1 = int(edges[1]) if e0 == int(edges[0]) else int(edges[0])
    g = float(net.edge_gamma[e0])
    g_new = g + window * (u - 0.5)
    while g_new <= 0.0 or g_new >= 1.0:
        if g_new <= 0.0:
            g_new = -g_new
        if g_new >= 1.0:
            g_new = 2.0 - g_new
    gam = net.edge_gamma.copy()
    gam[e0] = float(g_new)
    gam[e1] = float(1.0 - g_new)
    return net.replaced(edge_gamma=gam)


def gamma_move(net: Any, rng: np.random.Generator, window: float = 0.2) -> Optional[Tuple[Any, float]]:
    """``op_change_gamma`` without the graph (symmetric: log HR 0); ``None`` for a tree."""
    rets = reticulations(net)
    if rets.shape[0] == 0:
        return None
    r = int(rets[int(rng.integers(rets.shape[0]))])
    return gamma_move_at(net, r, float(rng.random()), window), 0.0
