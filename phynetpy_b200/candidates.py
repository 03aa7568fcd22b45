"""Candidate gene trees of the coupled moves, on flat arrays.

The reference's coupled add / delete / relocate-reticulation moves re-draw every locus' gene
tree from a finite set: the tree, its height-preserving NNI neighbours, each at the height scales
(0.25, 0.5, 1, 2, 4) (``_candidate_set`` ``src/_mcmc_seq.py:1836-1861``,
``_gene_tree_nni_neighbors`` ``:1724-1788``, ``_scaled_gene_tree`` ``:1819-1833``,
``_gt_signature`` ``:1699-1713``, ``_HEIGHT_SCALE_GRID`` ``:1816``) and score every member with a
full Felsenstein pruning (``:1935-1957``) -- up to ``5*(1+2(n-2))`` prunings per locus, forward and
reverse.  That loop is the dominant source of Felsenstein calls in a chain.

Here the set is enumerated on :class:`~phynetpy_b200.tree.FlatTree` arrays (no graph cloning) so
that ``SeqLikelihoodEngine.score_candidate_sets`` can score all loci x all candidates in one
device call.  Same semantics as the reference:

* node height = max over children of (child height + branch length), leaves at 0, a missing
  length counts as 0 (``src/GraphUtils.py:786-819``);
* an NNI around internal node ``v`` (parent ``p``, single sibling ``s``) swaps ``s`` with a child
  ``c`` of ``v`` and is allowed only if ``h[v] > h[s]``; node heights are preserved and every
  branch length is re-derived as ``max(h[parent] - h[child], 0)`` (``_sync_lengths`` ``:418-434``);
* scaling multiplies every node height by an exact power of two;
* candidates are de-duplicated by (descendant-leaf-set, round(height, 12)) of the internal nodes.
"""
from __future__ import annotations

from typing import FrozenSet, List, Sequence, Tuple

import numpy as np

from .tree import FlatTree

HEIGHT_SCALE_GRID: Tuple[float, ...] = (0.25, 0.5, 1.0, 2.0, 4.0)


def _children(parent: np.ndarray) -> List[List[int]]:
    kids: List[List[int]] = [[] for _ in range(parent.shape[0])]
    for i, p in enumerate(parent):
        if p >= 0:
            kids[int(p)].append(i)
    return kids


def _postorder(parent: np.ndarray, kids: List[List[int]]) -> List[int]:
    root = int(np.nonzero(parent < 0)[0][0])
    order, stack = [], [(root, 0)]
    while stack:
        v, k = stack.pop()
        if k < len(kids[v]):
            stack.append((v, k + 1))
            stack.append((kids[v][k], 0))
        else:
            order.append(v)
    return order


def node_heights(ft: FlatTree) -> np.ndarray:
    """Height of every node above the present (reference ``_node_heights``)."""
    kids = _children(ft.parent)
    h = np.zeros(ft.n_nodes, dtype=np.float64)
    blen = np.nan_to_num(ft.blen, nan=0.0)
    for v in _postorder(ft.parent, kids):
        best = 0.0
        for c in kids[v]:
            best = max(best, h[c] + float(blen[c]))
        h[v] = best
    return h


def _lengths_from_heights(parent: np.ndarray, h: np.ndarray) -> np.ndarray:
    out = np.full(parent.shape[0], np.nan, dtype=np.float64)
    nz = parent >= 0
    out[nz] = np.maximum(h[parent[nz]] - h[nz], 0.0)   # _sync_lengths clamps at 0
    return out


def _leaf_sets(ft: FlatTree, kids: List[List[int]]) -> List[FrozenSet]:
    sets: List[FrozenSet] = [frozenset()] * ft.n_nodes
    for v in _postorder(ft.parent, kids):
        if not kids[v]:
            sets[v] = frozenset([ft.label[v]])
        else:
            s = frozenset()
            for c in kids[v]:
                s = s | sets[c]
            sets[v] = s
    return sets


def signature(ft: FlatTree, heights: np.ndarray = None) -> FrozenSet:
    """Topology + time signature (reference ``_gt_signature``)."""
    kids = _children(ft.parent)
    h = node_heights(ft) if heights is None else heights
    ls = _leaf_sets(ft, kids)
    return frozenset((ls[v], round(float(h[v]), 12)) for v in range(ft.n_nodes) if kids[v])


def nni_neighbors(ft: FlatTree) -> List[Tuple[FlatTree, np.ndarray]]:
    """Height-preserving single-NNI rearrangements (reference ``_gene_tree_nni_neighbors``)."""
    return nni_neighbors_at(ft, node_heights(ft))


def nni_neighbors_at(ft: FlatTree, h: np.ndarray) -> List[Tuple[FlatTree, np.ndarray]]:
    """:func:`nni_neighbors` for a caller that already holds the node heights."""
    kids = _children(ft.parent)
    out, seen = [], set()
    for v in range(ft.n_nodes):
        if len(kids[v]) < 2 or ft.parent[v] < 0:
            continue
        p = int(ft.parent[v])
        sibs = [c for c in kids[p] if c != v]
        if len(sibs) != 1:
            continue
        s = sibs[0]
        if h[v] <= h[s]:
            continue
        for c in kids[v]:
            parent = ft.parent.copy()
            parent[s], parent[c] = v, p                    # s goes under v, c goes under p
            cand = FlatTree(parent, _lengths_from_heights(parent, h), ft.label)
            cand._tip_rows = ft._tip_rows   # an NNI re-hangs subtrees: leaves stay leaves, rows are unchanged
            sig = signature(cand, h)
            if sig in seen:
                continue
            seen.add(sig)
            out.append((cand, h))
    return out


def _candidate_arrays_plain(ft: FlatTree, scales: Sequence[float] = HEIGHT_SCALE_GRID) -> Tuple[np.ndarray, np.ndarray]:
    """Straightforward enumeration (one signature per member); kept as the cross-check of :func:`candidate_table`."""
    h0 = node_heights(ft)
    bases = [(ft, h0)] + nni_neighbors(ft)
    base_sigs = set()
    parents, lengths = [], []
    internal = ft.n_children() > 0
    for k, (base, h) in enumerate(bases):
        bsig = signature(base, h)
        if k > 0 and bsig in base_sigs:
            continue
        base_sigs.add(bsig)
        seen_h = set()
        for s in scales:
            hs = h * s if s != 1.0 else h
            key = tuple(np.round(hs[internal], 12).tolist())
            if key in seen_h:
                continue
            seen_h.add(key)
            parents.append(base.parent)
            lengths.append(ft.blen if (k == 0 and s == 1.0) else _lengths_from_heights(base.parent, hs))
    return np.ascontiguousarray(np.stack(parents), dtype=np.int32), np.ascontiguousarray(np.stack(lengths), dtype=np.float64)


def candidate_arrays(ft: FlatTree, scales: Sequence[float] = HEIGHT_SCALE_GRID) -> Tuple[np.ndarray, np.ndarray]:
    """The candidate set as stacked arrays ``(parent[K, n], blen[K, n])`` -- the form
    ``SeqLikelihoodEngine.score_candidate_sets`` packs without creating K tree objects.
    Same members and order as :func:`candidate_set`."""
    return candidate_table(ft, None, scales)[:2]


def clade_masks(parent: np.ndarray) -> List[int]:
    """Per node, the set of leaves below it as an integer bit set (bit = the leaf's node index); node indices are
    stable across NNI rearrangements of a flat tree, so masks of different members are comparable."""
    kids = _children(parent)
    clade = [0] * len(kids)
    for v in _postorder(parent, kids):
        if kids[v]:
            m = 0
            for c in kids[v]:
                m |= clade[c]
            clade[v] = m
        else:
            clade[v] = 1 << v
    return clade


def candidate_table(ft: FlatTree, heights: np.ndarray = None, scales: Sequence[float] = HEIGHT_SCALE_GRID,
                    info: dict = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """The candidate set with the node heights of every member: ``(parent[K, n], blen[K, n], height[K, n])`` --
    what a caller needs to score the members and to materialise the one it picks (``rebuild(k)`` of the reference,
    ``src/_mcmc_seq.py:1975-1986``).  ``heights`` are the tree's own node heights when the caller tracks them
    (the reference's ``gt_heights``); member (tree, scale 1) keeps the tree's lengths untouched (``:1828-1832``).

    Enumeration without per-member tree objects or signatures (the reference clones a graph per neighbour and
    hashes a set of leaf sets per member; on the host that -- not the scoring -- is what a coupled move costs once
    the scoring runs on the device):

    * an NNI around ``v`` changes exactly one clade (``v``'s) and no height, so neighbours are told apart by
      (v, new clade of v) with clades as integer bit sets;
    * the heights of a scaled member are the base's times the scale and its lengths are re-derived from them
      (``_sync_lengths`` on ``h * s``), vectorised over the scales;
    * two scales of one base can only collide after rounding to 12 decimals if every internal height is ~0;
      only then are the rounded heights compared.
    """
    h = node_heights(ft) if heights is None else np.ascontiguousarray(heights, dtype=np.float64)
    parent = ft.parent
    plist = parent.tolist()
    n = len(plist)
    kids: List[List[int]] = [[] for _ in range(n)]
    for i, p in enumerate(plist):
        if p >= 0:
            kids[p].append(i)
    # clades as bit sets, children before parents
    clade = [0] * n
    for v in _postorder(parent, kids):
        if kids[v]:
            m = 0
            for c in kids[v]:
                m |= clade[c]
            clade[v] = m
        else:
            clade[v] = 1 << v                      # bit = node index of the leaf (stable under NNI)
    hl = h.tolist()
    base_parents = [parent]
    base_change = [(-1, 0)]                        # per base: (node whose clade changed, its new clade)
    seen = set()
    for v in range(n):
        if len(kids[v]) < 2 or plist[v] < 0:
            continue
        p = plist[v]
        sibs = [c for c in kids[p] if c != v]
        if len(sibs) != 1:
            continue
        s = sibs[0]
        if hl[v] <= hl[s]:
            continue
        for c in kids[v]:
            key = (v, (clade[v] & ~clade[c]) | clade[s])
            if key in seen:
                continue
            seen.add(key)
            q = parent.copy()
            q[s], q[c] = v, p                      # s goes under v, c goes under p
            base_parents.append(q)
            base_change.append(key)
    sc = np.asarray(scales, dtype=np.float64)
    internal = np.fromiter((len(k) > 0 for k in kids), dtype=bool, count=n)
    hi = h[internal]
    # can two scales give the same rounded internal heights?  only if every internal height is tiny
    degenerate = hi.shape[0] == 0 or float(np.abs(hi).max()) * float(np.min(np.abs(np.diff(np.sort(sc))), initial=np.inf)) < 1e-9
    if degenerate:
        keep = []
        seen_h = set()
        for k, s_ in enumerate(sc.tolist()):
            key_h = tuple(np.round(hi * s_ if s_ != 1.0 else hi, 12).tolist())
            if key_h not in seen_h:
                seen_h.add(key_h)
                keep.append(k)
        sc = sc[keep]
    S = sc.shape[0]
    B = len(base_parents)
    P = np.repeat(np.stack(base_parents), S, axis=0)
    H = np.tile(h[None, :] * sc[:, None], (B, 1))
    H[np.tile(sc == 1.0, B)] = h                   # scale 1 is the height vector itself (no multiplication)
    L = np.full((B * S, n), np.nan, dtype=np.float64)
    Hs = H[:S]                                     # the S scaled height vectors (the same for every base)
    for b, q in enumerate(base_parents):           # _sync_lengths on the scaled heights, for any scale
        nz = q >= 0
        L[b * S:(b + 1) * S, nz] = np.maximum(Hs[:, q[nz]] - Hs[:, nz], 0.0)
    ident = np.nonzero(sc == 1.0)[0]
    if ident.shape[0]:
        L[int(ident[0])] = ft.blen                 # the tree itself at scale 1: lengths untouched
    if info is not None:     # what find_member needs: member b * S + k = base b at scale sc[k]
        info.update(clade=clade, internal=[v for v in range(n) if kids[v]], base_change=base_change, scales=sc.tolist(), heights=hl)
    return np.ascontiguousarray(P, dtype=np.int32), L, H


def find_member(info: dict, target_parent: np.ndarray, target_heights: np.ndarray) -> int:
    """Index of the member of a candidate table (built with ``info=``) that has the clades and the 12-decimal
    heights of the given tree -- the reference's ``_gt_signature`` match (``src/_mcmc_seq.py:2052-2058``) -- or -1.
    No member tree is built: base b differs from the table's tree in ONE clade, scale k multiplies the heights."""
    tm = clade_masks(target_parent)
    th = target_heights.tolist()
    t_kids = np.bincount(target_parent[target_parent >= 0], minlength=len(tm))
    target = {tm[u]: round(th[u], 12) for u in range(len(tm)) if t_kids[u] > 0}
    clade, internal, hl = info["clade"], info["internal"], info["heights"]
    S = len(info["scales"])
    for b, (v, new_mask) in enumerate(info["base_change"]):
        ok = True
        for u in internal:
            if (new_mask if u == v else clade[u]) not in target:
                ok = False
                break
        if not ok or len(internal) != len(target):
            continue
        for k, s_ in enumerate(info["scales"]):
            if all(round(hl[u] * s_ if s_ != 1.0 else hl[u], 12) == target[new_mask if u == v else clade[u]] for u in internal):
                return b * S + k
    return -1


def candidate_set(ft: FlatTree, scales: Sequence[float] = HEIGHT_SCALE_GRID) -> List[FlatTree]:
    """The re-proposal support of one locus: (tree + NNI neighbours) x height scales, de-duplicated
    (reference ``_candidate_set``).  The identity -- ``ft`` at scale 1 -- is always present."""
    h0 = node_heights(ft)
    bases = [(FlatTree(ft.parent, _lengths_from_heights(ft.parent, h0), ft.label), h0)]
    bases.extend(nni_neighbors(ft))
    out, seen = [], set()
    for k, (base, h) in enumerate(bases):
        for s in scales:
            hs = h * s if s != 1.0 else h
            # the reference leaves the tree itself untouched at scale 1 (:1828-1832)
            lengths = ft.blen if (k == 0 and s == 1.0) else _lengths_from_heights(base.parent, hs)
            cand = FlatTree.__new__(FlatTree)
            cand.parent, cand.blen, cand.label = base.parent, np.ascontiguousarray(lengths, dtype=np.float64), base.label
            cand._n_children = base._n_children
            cand._tip_rows = base._tip_rows if base._tip_rows is not None else ft._tip_rows
            sig = signature(cand, hs)
            if sig in seen:
                continue
            seen.add(sig)
            out.append(cand)
    return out


class CandidateBatch:
    """The candidate sets of MANY gene trees, enumerated natively (``pnb_candidates_count`` / ``_fill``) straight
    into the CSR arrays the device calls take.  Tree ``j`` owns members ``member_off[j] .. member_off[j+1]-1``;
    member ``m`` owns nodes ``node_off[m] .. node_off[m+1]-1`` of ``parent`` / ``blen`` / ``height``.
    Same members, order and bits as :func:`candidate_table` per tree."""

    __slots__ = ("trees", "member_off", "node_off", "parent", "blen", "height", "changed_node", "scale", "tree_of_member")

    def __init__(self, trees: Sequence[FlatTree], heights: Sequence[np.ndarray], scales: Sequence[float] = HEIGHT_SCALE_GRID,
                 context=None):
        from . import _lib
        lib = _lib.load()
        self.trees = list(trees)
        T = len(self.trees)
        sizes = np.fromiter((t.n_nodes for t in self.trees), dtype=np.int64, count=T)
        tree_off = np.zeros(T + 1, dtype=np.int64)
        np.cumsum(sizes, out=tree_off[1:])
        parent = np.ascontiguousarray(np.concatenate([t.parent for t in self.trees]) if T else np.zeros(0), dtype=np.int32)
        blen = np.ascontiguousarray(np.concatenate([t.blen for t in self.trees]) if T else np.zeros(0), dtype=np.float64)
        h = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.float64) for x in heights]) if T else np.zeros(0),
                                 dtype=np.float64)
        if h.shape[0] != parent.shape[0]:
            raise ValueError("CandidateBatch: one height per node of every tree")
        sc = np.ascontiguousarray(scales, dtype=np.float64)
        ctx_h = context.handle if context is not None else None
        counts = np.zeros(T, dtype=np.int64)
        _lib.check(lib.pnb_candidates_count(ctx_h, T, _lib.ptr(tree_off), _lib.ptr(parent), _lib.ptr(h), int(sc.shape[0]),
                                            _lib.ptr(sc), _lib.ptr(counts)))
        self.member_off = np.zeros(T + 1, dtype=np.int64)
        np.cumsum(counts, out=self.member_off[1:])
        n_members = int(self.member_off[-1])
        member_sizes = np.repeat(sizes, counts)
        self.node_off = np.zeros(n_members + 1, dtype=np.int64)
        np.cumsum(member_sizes, out=self.node_off[1:])
        rows = int(self.node_off[-1])
        self.parent = np.empty(rows, dtype=np.int32)
        self.blen = np.empty(rows, dtype=np.float64)
        self.height = np.empty(rows, dtype=np.float64)
        self.changed_node = np.empty(n_members, dtype=np.int32)
        self.scale = np.empty(n_members, dtype=np.float64)
        self.tree_of_member = np.repeat(np.arange(T, dtype=np.int64), counts)
        _lib.check(lib.pnb_candidates_fill(ctx_h, T, _lib.ptr(tree_off), _lib.ptr(parent), _lib.ptr(blen), _lib.ptr(h),
                                           int(sc.shape[0]), _lib.ptr(sc), _lib.ptr(self.member_off), _lib.ptr(self.parent),
                                           _lib.ptr(self.blen), _lib.ptr(self.height), _lib.ptr(self.changed_node),
                                           _lib.ptr(self.scale)))

    @property
    def n_members(self) -> int:
        return int(self.member_off[-1])

    def expand(self, per_tree: Sequence[np.ndarray], dtype) -> np.ndarray:
        """One per-node array per TREE (tip rows, species ids) repeated for every member of that tree."""
        counts = np.diff(self.member_off)
        sizes = {a.shape[0] for a in per_tree}
        if len(sizes) == 1 and len(per_tree):       # equal tree sizes: one repeat over a 2-D array
            return np.ascontiguousarray(np.repeat(np.stack(per_tree), counts, axis=0).reshape(-1), dtype=dtype)
        return np.ascontiguousarray(np.concatenate([np.tile(a, int(k)) for a, k in zip(per_tree, counts)]), dtype=dtype)

    def member(self, m: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        a, b = int(self.node_off[m]), int(self.node_off[m + 1])
        return self.parent[a:b], self.blen[a:b], self.height[a:b]

    def table(self, j: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Members of tree ``j`` as ``(parent[K, n], blen[K, n], height[K, n])`` views (the form of candidate_table)."""
        m0, m1 = int(self.member_off[j]), int(self.member_off[j + 1])
        a, b = int(self.node_off[m0]), int(self.node_off[m1])
        n = self.trees[j].n_nodes
        return self.parent[a:b].reshape(m1 - m0, n), self.blen[a:b].reshape(m1 - m0, n), self.height[a:b].reshape(m1 - m0, n)


def find_member_in_batch(batch: CandidateBatch, j: int, target_parent: np.ndarray, target_heights: np.ndarray) -> int:
    """:func:`find_member` for tree ``j`` of a natively enumerated batch: index (within that tree's members) of the
    member with the clades and 12-decimal heights of the given tree, or -1.  Uses what the enumerator recorded per
    member -- the node whose clade the NNI changed and the scale -- so no member tree is rebuilt."""
    ft = batch.trees[j]
    m0, m1 = int(batch.member_off[j]), int(batch.member_off[j + 1])
    # common case first: node identities are shared by a tree and everything enumerated from it, so the original
    # usually IS a member parent array for parent array (an NNI undone) with node-wise equal rounded heights
    if target_parent.shape[0] == ft.n_nodes:
        P, _L, H = batch.table(j)
        inner = ft.n_children() > 0
        hit = np.nonzero((P == target_parent[None, :]).all(axis=1)
                         & (np.round(H[:, inner], 12) == np.round(target_heights[inner], 12)[None, :]).all(axis=1))[0]
        if hit.shape[0]:
            return int(hit[0])
    clade = clade_masks(ft.parent)
    tm = clade_masks(target_parent)
    th = target_heights.tolist()
    t_kids = np.bincount(target_parent[target_parent >= 0], minlength=len(tm))
    target = {tm[u]: round(th[u], 12) for u in range(len(tm)) if t_kids[u] > 0}
    internal = np.nonzero(ft.n_children() > 0)[0].tolist()
    if len(internal) != len(target):
        return -1
    topo_ok = {}                      # (changed node, its new clade) -> the member's clades are the target's
    for m in range(m0, m1):
        v = int(batch.changed_node[m])
        par, _bl, ht = batch.member(m)
        new_mask = 0
        if v >= 0:                    # the member's clades are the tree's, except node v's
            for c in np.nonzero(par == v)[0].tolist():
                new_mask |= clade[c]
        ok = topo_ok.get((v, new_mask))
        if ok is None:
            ok = all((new_mask if u == v else clade[u]) in target for u in internal)
            topo_ok[(v, new_mask)] = ok
        if not ok:
            continue
        hl = ht.tolist()
        if all(round(hl[u], 12) == target[new_mask if u == v else clade[u]] for u in internal):
            return m - m0
    return -1
