"""Starting state of an MCMC_SEQ chain on flat objects: UPGMA gene trees from the alignments and the shrinking of
the species network that makes every starting gene tree embed.

* :func:`jc_distance_matrix` / :func:`build_upgma_gene_tree` -- the reference's ``_jc_distance`` and
  ``build_upgma_gene_tree`` (``src/_mcmc_seq.py:548-641``; PhyloNet's default ``-sgt``): Jukes-Cantor distances over
  the columns where both sequences are one of A, C, G, T (no such column -> 1.0; p >= 3/4 -> 3.0), average-linkage
  clustering that always merges the first closest pair in its scan order, node heights = half the cluster distance,
  branch lengths floored at 1e-9 and -- because the reference assembles a Newick string with ``:.10f`` and parses it
  back -- rounded to ten decimals.  The distance matrix is one vectorised comparison of the byte matrix (all pairs x
  all sites) instead of a Python loop per pair and per character;
* :func:`enforce_embedding_consistency` -- ``_enforce_embedding_consistency`` (``:473-541``): every cross-species
  coalescence bounds the height of the species node that is the MRCA of its species; the network is scaled by the
  tightest ``(1 - eps) * bound / height`` below 1.

Host code (setup, once per chain).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np

from ._msnc_density import SpeciesNetwork
from .tree import FlatTree

_ACGT = np.zeros(256, dtype=bool)
for _c in b"ACGTacgt":
    _ACGT[_c] = True
_UPPER = np.arange(256, dtype=np.uint8)
for _c in range(ord("a"), ord("z") + 1):
    _UPPER[_c] = _c - 32


def jc_distance_matrix(alignment: Dict[str, str]) -> Tuple[List[str], np.ndarray]:
    """Labels (dict order) and the symmetric matrix of Jukes-Cantor distances (``_jc_distance``)."""
    labels = list(alignment.keys())
    n = len(labels)
    if n == 0:
        return labels, np.zeros((0, 0))
    L = min(len(alignment[lab]) for lab in labels)          # zip() in the reference stops at the shorter sequence
    mat = np.frombuffer("".join(alignment[lab][:L] for lab in labels).encode("latin-1", "replace"), dtype=np.uint8).reshape(n, L)
    up = _UPPER[mat]
    ok = _ACGT[mat]
    both = ok[:, None, :] & ok[None, :, :]
    n_cmp = both.sum(axis=2)
    n_diff = (both & (up[:, None, :] != up[None, :, :])).sum(axis=2)
    d = np.ones((n, n), dtype=np.float64)
    for i in range(n):
        for j in range(n):
            if n_cmp[i, j] == 0:
                d[i, j] = 1.0
                continue
            p = int(n_diff[i, j]) / int(n_cmp[i, j])
            d[i, j] = 3.0 if p >= 0.75 else -0.75 * math.log(1.0 - (4.0 / 3.0) * p)
    return labels, d


def build_upgma_gene_tree(alignment: Dict[str, str]) -> FlatTree:
    """The reference's UPGMA starting gene tree of one locus as a :class:`FlatTree` (leaves first, in dict order;
    internal nodes ``I<n>`` in the order they are created; the root last)."""
    labels, d0 = jc_distance_matrix(alignment)
    n = len(labels)
    if n == 0:
        raise ValueError("build_upgma_gene_tree: empty alignment")
    if n == 1:
        return FlatTree([-1], [None], [labels[0]])
    dist: Dict[Tuple[int, int], float] = {(i, j): float(d0[i, j]) for i in range(n) for j in range(i + 1, n)}
    height = [0.0] * n
    size = [1] * n
    parent = [-1] * n
    blen: List[float] = [math.nan] * n
    label: List = list(labels)
    active = list(range(n))
    while len(active) > 1:
        best, best_d = None, math.inf
        for a in range(len(active)):
            for b in range(a + 1, len(active)):
                ia, ib = active[a], active[b]
                dd = dist[(min(ia, ib), max(ia, ib))]
                if dd < best_d:
                    best_d, best = dd, (a, b)
        a, b = best
        ia, ib = active[a], active[b]
        new_h = best_d / 2.0
        new = len(height)
        for child in (ia, ib):
            parent[child] = new
            blen[child] = float("%.10f" % max(new_h - height[child], 1e-9))    # through the Newick text, as the reference
        height.append(new_h)
        size.append(size[ia] + size[ib])
        parent.append(-1)
        blen.append(math.nan)
        label.append(None)
        remaining = [x for k, x in enumerate(active) if k not in (a, b)]
        for ic in remaining:
            d_ac = dist[(min(ia, ic), max(ia, ic))]
            d_bc = dist[(min(ib, ic), max(ib, ic))]
            dist[(min(new, ic), max(new, ic))] = (size[ia] * d_ac + size[ib] * d_bc) / (size[ia] + size[ib])
        active = remaining + [new]
    return FlatTree(parent, [None if x != x else x for x in blen], label)


def _species_sets_tree(ft: FlatTree, species_of: Dict[str, str]) -> List[frozenset]:
    n = ft.n_nodes
    kids: List[List[int]] = [[] for _ in range(n)]
    for i, p in enumerate(ft.parent.tolist()):
        if p >= 0:
            kids[p].append(i)
    out: List = [None] * n

    def desc(v: int) -> frozenset:
        stack = [v]
        while stack:
            x = stack[-1]
            if out[x] is not None:
                stack.pop()
                continue
            pend = [c for c in kids[x] if out[c] is None]
            if pend:
                stack.extend(pend)
                continue
            out[x] = frozenset({species_of.get(ft.label[x], ft.label[x])}) if not kids[x] else frozenset().union(*[out[c] for c in kids[x]])
            stack.pop()
        return out[v]
    for v in range(n):
        desc(v)
    return out


def _species_sets_net(net: SpeciesNetwork) -> List[frozenset]:
    n = net.n_nodes
    kids: List[List[int]] = [[] for _ in range(n)]
    for s, d in zip(net.edge_src.tolist(), net.edge_dst.tolist()):
        kids[s].append(d)
    out: List = [None] * n
    order = np.argsort(net.height, kind="stable").tolist()
    for _ in range(n + 1):
        done = True
        for v in order:
            if out[v] is not None:
                continue
            if not kids[v]:
                out[v] = frozenset({net.node_label[v]})
            elif all(out[c] is not None for c in kids[v]):
                out[v] = frozenset().union(*[out[c] for c in kids[v]])
            else:
                done = False
        if done:
            break
    return out


def enforce_embedding_consistency(net: SpeciesNetwork, gene_trees: Sequence[FlatTree], heights: Sequence[np.ndarray],
                                  species_of: Dict[str, str], eps: float = 0.05) -> Tuple[SpeciesNetwork, float]:
    """Scale the species network down until every gene tree embeds (``_enforce_embedding_consistency``): returns the
    (possibly scaled) network and the scale factor, 1.0 if nothing had to change."""
    sp_desc = _species_sets_net(net)
    h = net.height
    nodes_by_h = sorted(range(net.n_nodes), key=lambda v: h[v])
    bounds = [math.inf] * net.n_nodes
    for ft, gh in zip(gene_trees, heights):
        gdesc = _species_sets_tree(ft, species_of)
        nk = ft.n_children()
        for u in range(ft.n_nodes):
            if nk[u] == 0 or len(gdesc[u]) < 2:
                continue
            hu = float(gh[u])
            for v in nodes_by_h:                      # the lowest node covering the set = the MRCA
                if gdesc[u] <= sp_desc[v]:
                    if hu < bounds[v]:
                        bounds[v] = hu
                    break
    scale = 1.0
    for v, upper in enumerate(bounds):
        if math.isfinite(upper) and h[v] > 0.0:
            ratio = (1.0 - eps) * upper / float(h[v])
            if ratio < scale:
                scale = ratio
    if scale < 1.0:
        return net.replaced(height=net.height * scale), scale
    return net, 1.0
