"""Reticulation surgery and the placement scan of the informed add / delete moves, on flat species networks.

The reference's ``op_add_reticulation`` / ``op_delete_reticulation`` (``src/_mcmc_seq.py:1365-1625``) rank every
ordered edge pair (donor, reticulation edge) by the network-only score of a representative placement
(``_add_reticulation_placements`` ``:1232-1318`` over ``_network_only_log_score`` ``:1163-1217``): for E edges that is
~E^2 candidate networks, each scored by ``M loci x 5 height scales`` MSNC densities on the host -- by far the most
expensive thing a reticulation move does.  Here

* a candidate network is a handful of array edits (:func:`insert_reticulation`; no graph cloning),
* its score is ONE ``pnb_msnc_eval`` over all loci x all scales (:func:`network_only_msnc_score`),

so the scan is E^2 device calls of a few hundred microseconds instead of ``5 M E^2`` host evaluations.  The prior
(``log_prior_seq``, sampler side) is the caller's: pass it as ``prior_fn`` and it is added to every score exactly
where the reference adds it.

Edges are identified by index into the network's edge arrays; new nodes are appended (the donor point first, the
hybrid node second), new and split edges are appended / rewritten so that every untouched edge keeps its index.
A split resets the inheritance probability of the edge it splits, as the reference's ``_split_edge`` does
(``:1116-1129``): its two halves are fresh ``Edge`` objects, whose gamma defaults to 0.0 (``src/Network.py:696-699``).
For a half that enters a tree node that value is never read; for the lower half of an edge that entered an
EXISTING reticulation it means that parent now carries probability 0 (log-floor in the density), which is what the
reference scores -- reproduced, not repaired.
"""
from __future__ import annotations

import math
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from ._msnc_density import SpeciesNetwork, msnc_log_density_batch
from .candidates import HEIGHT_SCALE_GRID
from .tree import FlatTree


def has_parallel_edges(net: SpeciesNetwork) -> bool:
    """Two edges between the same ordered pair of nodes -- a "bubble" (``_has_parallel_edges`` ``:1132-1147``)."""
    pairs = net.edge_src.astype(np.int64) * net.n_nodes + net.edge_dst
    return np.unique(pairs).shape[0] != pairs.shape[0]


def is_acyclic(net: SpeciesNetwork) -> bool:
    n = net.n_nodes
    indeg = np.bincount(net.edge_dst, minlength=n).tolist()
    out: List[List[int]] = [[] for _ in range(n)]
    for s, d in zip(net.edge_src.tolist(), net.edge_dst.tolist()):
        out[s].append(d)
    ready = [v for v in range(n) if indeg[v] == 0]
    seen = 0
    while ready:
        v = ready.pop()
        seen += 1
        for c in out[v]:
            indeg[c] -= 1
            if indeg[c] == 0:
                ready.append(c)
    return seen == n


def _unique_label(net: SpeciesNetwork, stem: str) -> str:
    used = set(x for x in net.node_label if x is not None)
    k = 0
    while "%s%d" % (stem, k) in used:
        k += 1
    return "%s%d" % (stem, k)


def insert_reticulation(net: SpeciesNetwork, donor_edge: int, retic_edge: int, t1: float, t2: float,
                        gamma: float) -> Optional[SpeciesNetwork]:
    """The surgery of ``op_add_reticulation`` (``:1433-1452``): a tree node ``v1`` at height ``t1`` on the donor edge,
    a hybrid node ``v2`` at height ``t2`` on the reticulation edge, a hybrid edge ``v1 -> v2`` with ``gamma``; the
    edge entering ``v2`` from above gets ``1 - gamma``.  ``None`` if the result has a bubble or a cycle."""
    if donor_edge == retic_edge:
        return None
    n = net.n_nodes
    src, dst, gam = net.edge_src.tolist(), net.edge_dst.tolist(), net.edge_gamma.tolist()
    v3, v4 = src[donor_edge], dst[donor_edge]
    v5, v6 = src[retic_edge], dst[retic_edge]
    v1, v2 = n, n + 1
    # donor edge v3 -> v4 becomes v3 -> v1 (in place) and v1 -> v4 (appended); fresh edges: gamma 0.0
    src[donor_edge], dst[donor_edge], gam[donor_edge] = v3, v1, 0.0
    src.append(v1); dst.append(v4); gam.append(0.0)
    # reticulation edge v5 -> v6 becomes v5 -> v2 (in place, 1 - gamma) and v2 -> v6 (appended, fresh)
    src[retic_edge], dst[retic_edge], gam[retic_edge] = v5, v2, 1.0 - gamma
    src.append(v2); dst.append(v6); gam.append(0.0)
    # the hybrid edge
    src.append(v1); dst.append(v2); gam.append(float(gamma))
    height = np.concatenate([net.height, [t1, t2]])
    node_label = list(net.node_label) + [_unique_label(net, "x_"), _unique_label(net, "#H_")]
    leaf_label = list(net.leaf_label) + [None, None]
    out = SpeciesNetwork(src, dst, [None if g != g else g for g in gam], height, leaf_label, node_label)
    if has_parallel_edges(out) or not is_acyclic(out):
        return None
    return out


def remove_reticulation(net: SpeciesNetwork, retic: int, donor_parent: int) -> Optional[SpeciesNetwork]:
    """The surgery of ``op_delete_reticulation`` (``:1545-1567``): the hybrid edge ``donor_parent -> retic`` is
    removed and the two resulting degree-2 nodes are suppressed (``_suppress_degree2`` ``:1628-1655``: the merged
    edge keeps the gamma of its LOWER half when its child is a reticulation).  ``None`` for the configurations the
    reference declines (``:1517-1537``)."""
    src, dst, gam = net.edge_src.tolist(), net.edge_dst.tolist(), net.edge_gamma.tolist()
    n = net.n_nodes
    in_e = [[] for _ in range(n)]
    out_e = [[] for _ in range(n)]
    for e, (s, d) in enumerate(zip(src, dst)):
        out_e[s].append(e)
        in_e[d].append(e)
    v2, v1 = retic, donor_parent
    parents = [src[e] for e in in_e[v2]]
    if len(parents) != 2 or v1 not in parents:
        return None
    v5 = parents[1] if parents[0] == v1 else parents[0]
    if len(in_e[v1]) != 1:                       # v1 must be a tree node below the root
        return None
    v3 = src[in_e[v1][0]]
    others = [dst[e] for e in out_e[v1] if dst[e] != v2]
    if len(others) != 1 or len(out_e[v2]) != 1:
        return None
    v4, v6 = others[0], dst[out_e[v2][0]]
    if v3 in [src[e] for e in in_e[v4]] or v5 in [src[e] for e in in_e[v6]] or (v3 == v5 and v4 == v6):
        return None
    is_retic = [len(x) >= 2 for x in in_e]
    edges = [(s, d, g) for e, (s, d, g) in enumerate(zip(src, dst, gam)) if not (s == v1 and d == v2)]

    def suppress(edges, node):
        ins = [k for k, (s, d, g) in enumerate(edges) if d == node]
        outs = [k for k, (s, d, g) in enumerate(edges) if s == node]
        if len(ins) != 1 or len(outs) != 1:
            return edges, False
        p, c, g_out = edges[ins[0]][0], edges[outs[0]][1], edges[outs[0]][2]
        child_is_retic = is_retic[c] and c != v2     # the node FLAG of the reference; v2's was just cleared
        merged = (p, c, g_out if child_is_retic else 0.0)   # a fresh Edge: gamma 0.0 unless it enters a reticulation
        keep = [x for k, x in enumerate(edges) if k not in (ins[0], outs[0])]
        return keep + [merged], True

    edges, ok1 = suppress(edges, v1)
    edges, ok2 = suppress(edges, v2)
    gone = {v for v, ok in ((v1, ok1), (v2, ok2)) if ok}
    if len(edges) < 2:
        return None
    keep_nodes = [v for v in range(n) if v not in gone]
    remap = {old: new for new, old in enumerate(keep_nodes)}
    return SpeciesNetwork([remap[s] for s, d, g in edges], [remap[d] for s, d, g in edges],
                          [None if g != g else g for s, d, g in edges], net.height[keep_nodes],
                          [net.leaf_label[v] for v in keep_nodes], [net.node_label[v] for v in keep_nodes])


def geometric_placements(net: SpeciesNetwork) -> List[Tuple[int, int]]:
    """Every ordered edge pair (donor, reticulation edge) with positive lengths and a non-empty height window --
    ``_geometric_placements`` (``:1321-1362``); feasibility of the surgery is not checked here."""
    h = net.height
    src, dst = net.edge_src, net.edge_dst
    out = []
    for d in range(src.shape[0]):
        if h[src[d]] - h[dst[d]] <= 0.0:
            continue
        for r in range(src.shape[0]):
            if r == d or h[src[r]] - h[dst[r]] <= 0.0:
                continue
            if h[src[d]] <= max(h[dst[d]], h[dst[r]]):
                continue
            out.append((d, r))
    return out


def representative_placement(net: SpeciesNetwork, donor_edge: int, retic_edge: int) -> Optional[Tuple[SpeciesNetwork, float]]:
    """The midpoint placement the scan scores for an edge pair (``:1268-1303``): hybrid node at the middle of the
    reticulation edge, donor point at the middle of the feasible window, gamma 1/2.  Returns ``(network, l2)``."""
    h = net.height
    v3, v4 = int(net.edge_src[donor_edge]), int(net.edge_dst[donor_edge])
    v5, v6 = int(net.edge_src[retic_edge]), int(net.edge_dst[retic_edge])
    if donor_edge == retic_edge or h[v3] - h[v4] <= 0.0:
        return None
    l2 = float(h[v5] - h[v6])
    if l2 <= 0.0:
        return None
    t2 = float(h[v6]) + 0.5 * l2
    lo, hi = max(float(h[v4]), t2), float(h[v3])
    if hi <= lo:
        return None
    cand = insert_reticulation(net, donor_edge, retic_edge, 0.5 * (lo + hi), t2, 0.5)
    return None if cand is None else (cand, l2)


def scaled_gene_trees(trees: Sequence[FlatTree], scales: Sequence[float] = HEIGHT_SCALE_GRID) -> List[FlatTree]:
    """Every gene tree at every height scale, tree-major (tree 0 at all scales, then tree 1, ...): the reference
    scales the coalescence TIMES (``:1196-1199``); multiplying the branch lengths by an exact power of two gives the
    same times bit for bit."""
    out = []
    for t in trees:
        for s in scales:
            out.append(t if s == 1.0 else t.with_lengths(t.blen * s))
    return out


def network_only_msnc_score(per_tree_scale: np.ndarray, n_scales: int) -> float:
    """``sum_i logsumexp_s log P(g_i scaled by s | net)`` from the densities of :func:`scaled_gene_trees`
    (``_network_only_log_score`` ``:1186-1214`` without the prior); ``-inf`` if some locus embeds at no scale."""
    d = np.asarray(per_tree_scale, dtype=np.float64).reshape(-1, n_scales)
    total = 0.0
    for row in d.tolist():
        terms = [x for x in row if math.isfinite(x)]
        if not terms:
            return -math.inf
        m = max(terms)
        acc = 0.0
        for v in terms:
            acc += math.exp(v - m)
        total += m + math.log(acc)
    return total


def add_reticulation_placements(net: SpeciesNetwork, density_fn: Callable[[SpeciesNetwork], np.ndarray], n_scales: int,
                                prior_fn: Optional[Callable[[SpeciesNetwork], float]] = None) -> List[Dict[str, Any]]:
    """The scored placement list of the informed add / delete moves (``_add_reticulation_placements``).

    ``density_fn(candidate_network)`` returns the MSNC densities of :func:`scaled_gene_trees` under the candidate --
    e.g. ``lambda n: msnc_log_density_batch(scaled, n, species_of, theta)``, one device call; ``prior_fn`` is the
    sampler's ``log_prior_seq`` on the candidate (0 if omitted).  Returns dicts with ``donor_edge``, ``retic_edge``,
    ``l2`` and ``s_rep``, in the reference's order (donor edge outer, reticulation edge inner); placements whose
    surgery fails or whose score is not finite are dropped, as there."""
    out = []
    E = int(net.edge_src.shape[0])
    for d in range(E):
        for r in range(E):
            rep = representative_placement(net, d, r)
            if rep is None:
                continue
            cand, l2 = rep
            s = network_only_msnc_score(density_fn(cand), n_scales)
            if prior_fn is not None and math.isfinite(s):
                s += float(prior_fn(cand))
            if not math.isfinite(s):
                continue
            out.append({"donor_edge": d, "retic_edge": r, "l2": l2, "s_rep": s})
    return out


def device_density_fn(trees: Sequence[FlatTree], species_of: Dict[str, str], theta: float,
                      scales: Sequence[float] = HEIGHT_SCALE_GRID, context: Any = None,
                      branch_thetas: Optional[dict] = None) -> Callable[[SpeciesNetwork], np.ndarray]:
    """The ``density_fn`` of :func:`add_reticulation_placements` backed by the device: the gene trees are scaled
    once; every candidate network then costs one ``pnb_msnc_eval`` over ``len(trees) * len(scales)`` trees."""
    scaled = scaled_gene_trees(trees, scales)

    def fn(cand: SpeciesNetwork) -> np.ndarray:
        return msnc_log_density_batch(scaled, cand, species_of, theta, branch_thetas=branch_thetas, context=context)
    return fn


# ---------------------------------------------------------------------------------------------------
# the informed add / delete moves themselves (network part; src/_mcmc_seq.py:1365-1625)
# ---------------------------------------------------------------------------------------------------
def _logsumexp(values: Sequence[float]) -> float:
    m = -math.inf
    for v in values:
        if v > m:
            m = v
    if m == -math.inf:
        return -math.inf
    acc = 0.0
    for v in values:
        acc += math.exp(v - m)
    return m + math.log(acc)


def placement_probabilities(placements: Sequence[Dict[str, Any]]) -> Tuple[np.ndarray, float]:
    """Selection probabilities ``exp(s_rep - logsumexp)`` (renormalised) and the log-sum-exp (``:1411-1415``)."""
    s = [p["s_rep"] for p in placements]
    lse = _logsumexp(s)
    probs = np.asarray([math.exp(x - lse) for x in s], dtype=np.float64)
    probs /= probs.sum()
    return probs, lse


def add_reticulation_at(net: SpeciesNetwork, placements: Sequence[Dict[str, Any]], k: int, u_t2: float, u_t1: float,
                        gamma: float) -> Optional[Tuple[SpeciesNetwork, float]]:
    """The deterministic half of ``op_add_reticulation``: placement ``k`` with the uniforms for the hybrid height,
    the donor height and gamma.  Returns ``(network, log HR)`` with

        log HR = logsumexp_k s_rep_k - s_rep* - log(2 (R + 1)) + log l2 + log win        (``:1454-1461``)

    or ``None`` where the reference declines (empty window, bubble, cycle)."""
    _, lse = placement_probabilities(placements)
    chosen = placements[k]
    d, r = int(chosen["donor_edge"]), int(chosen["retic_edge"])
    h = net.height
    v3, v4 = int(net.edge_src[d]), int(net.edge_dst[d])
    v5, v6 = int(net.edge_src[r]), int(net.edge_dst[r])
    l2 = float(h[v5] - h[v6])
    if l2 <= 0.0:
        return None
    t2 = float(h[v6]) + l2 * u_t2
    lo, hi = max(float(h[v4]), t2), float(h[v3])
    win = hi - lo
    if win <= 0.0:
        return None
    t1 = lo + win * u_t1
    new = insert_reticulation(net, d, r, t1, t2, float(gamma))
    if new is None:
        return None
    R = net.n_reticulations()
    loghr = lse - chosen["s_rep"] - math.log(2.0 * (R + 1)) + math.log(l2) + math.log(win)
    return new, loghr


def add_reticulation_move(net: SpeciesNetwork, placements: Sequence[Dict[str, Any]], rng: np.random.Generator
                          ) -> Optional[Tuple[SpeciesNetwork, float]]:
    """``op_add_reticulation`` without the graph; draws in the reference's order (placement, t2, t1, gamma)."""
    if not placements:
        return None
    probs, _ = placement_probabilities(placements)
    k = int(rng.choice(len(placements), p=probs))
    u2 = float(rng.random())
    # the reference stops before drawing t1 / gamma when the window is empty (``:1426-1431``)
    chosen = placements[k]
    d, r = int(chosen["donor_edge"]), int(chosen["retic_edge"])
    h = net.height
    l2 = float(h[net.edge_src[r]] - h[net.edge_dst[r]])
    if l2 <= 0.0:
        return None
    t2 = float(h[net.edge_dst[r]]) + l2 * u2
    if float(h[net.edge_src[d]]) - max(float(h[net.edge_dst[d]]), t2) <= 0.0:
        return None
    u1 = float(rng.random())
    g = float(rng.random())
    return add_reticulation_at(net, placements, k, u2, u1, g)


def delete_reticulation_geometry(net: SpeciesNetwork, retic: int, donor_parent: int) -> Optional[Dict[str, Any]]:
    """The network half of a delete: the post-delete network plus what the reverse add needs -- the labels of the
    merged donor and reticulation edges, ``l2`` and the donor window (``:1538-1551``).  ``None`` where the reference
    declines."""
    src, dst = net.edge_src.tolist(), net.edge_dst.tolist()
    parents = [src[e] for e in range(len(src)) if dst[e] == retic]
    if len(parents) != 2 or donor_parent not in parents:
        return None
    v1, v2 = donor_parent, retic
    v5 = parents[1] if parents[0] == v1 else parents[0]
    up1 = [src[e] for e in range(len(src)) if dst[e] == v1]
    if len(up1) != 1:
        return None
    v3 = up1[0]
    others = [dst[e] for e in range(len(src)) if src[e] == v1 and dst[e] != v2]
    down2 = [dst[e] for e in range(len(src)) if src[e] == v2]
    if len(others) != 1 or len(down2) != 1:
        return None
    v4, v6 = others[0], down2[0]
    h = net.height
    l2 = float(h[v5] - h[v6])
    if l2 <= 0.0:
        return None
    lo, hi = max(float(h[v4]), float(h[v2])), float(h[v3])
    win = hi - lo
    if win <= 0.0:
        return None
    new = remove_reticulation(net, v2, v1)
    if new is None:
        return None
    lab = net.node_label
    return {"network": new, "donor_labels": (lab[v3], lab[v4]), "retic_labels": (lab[v5], lab[v6]), "l2": l2, "win": win,
            "R": net.n_reticulations()}


def delete_reticulation_loghr(geom: Dict[str, Any], placements: Sequence[Dict[str, Any]]) -> Optional[float]:
    """``s_rep(reverse placement) - logsumexp_k s_rep_k + log(2 R) - log l2 - log win`` (``:1597-1603``) from the
    placements of the post-delete network; ``None`` if the reverse add cannot recreate the removed reticulation."""
    if not placements:
        return None
    new = geom["network"]
    lab = new.node_label
    _, lse = placement_probabilities(placements)
    for p in placements:
        d, r = int(p["donor_edge"]), int(p["retic_edge"])
        if (lab[int(new.edge_src[d])], lab[int(new.edge_dst[d])]) == geom["donor_labels"] and \
           (lab[int(new.edge_src[r])], lab[int(new.edge_dst[r])]) == geom["retic_labels"]:
            return p["s_rep"] - lse + math.log(2.0 * geom["R"]) - math.log(geom["l2"]) - math.log(geom["win"])
    return None


def delete_reticulation_at(net: SpeciesNetwork, retic: int, donor_parent: int,
                           placements_fn: Callable[[SpeciesNetwork], List[Dict[str, Any]]]
                           ) -> Optional[Tuple[SpeciesNetwork, float]]:
    """The deterministic half of ``op_delete_reticulation``: remove the hybrid edge ``donor_parent -> retic``, score
    the placements of the resulting network with the add move's own rule (``placements_fn``) and return
    ``(network, log HR)``; ``None`` where the reference declines."""
    geom = delete_reticulation_geometry(net, retic, donor_parent)
    if geom is None:
        return None
    loghr = delete_reticulation_loghr(geom, placements_fn(geom["network"]))
    return None if loghr is None else (geom["network"], loghr)


def draw_reticulation_to_delete(net: SpeciesNetwork, rng: np.random.Generator) -> Optional[Tuple[int, int]]:
    """A reticulation uniformly, then one of its two parents uniformly (``:1500-1506``)."""
    rets = np.nonzero(np.bincount(net.edge_dst, minlength=net.n_nodes) >= 2)[0]
    if rets.shape[0] == 0:
        return None
    v2 = int(rets[int(rng.integers(rets.shape[0]))])
    parents = net.edge_src[net.edge_dst == v2]
    if parents.shape[0] != 2:
        return None
    return v2, int(parents[int(rng.integers(2))])


def delete_reticulation_move(net: SpeciesNetwork, rng: np.random.Generator,
                             placements_fn: Callable[[SpeciesNetwork], List[Dict[str, Any]]]
                             ) -> Optional[Tuple[SpeciesNetwork, float]]:
    """``op_delete_reticulation`` without the graph."""
    pick = draw_reticulation_to_delete(net, rng)
    return None if pick is None else delete_reticulation_at(net, pick[0], pick[1], placements_fn)


def network_level(net: SpeciesNetwork) -> int:
    """The level of a network: the largest number of reticulation nodes in one biconnected component of its
    undirected graph, 0 for a tree (``level`` / ``blobs``, ``src/GraphUtils.py:603-732``; Tarjan's algorithm, here
    without recursion).  The reticulation moves reject proposals above ``priors.max_level`` before paying for the
    gene-tree re-proposal (``src/_mcmc_seq.py:2150-2155``)."""
    n = net.n_nodes
    if n == 0:
        return 0
    adj: List[List[int]] = [[] for _ in range(n)]
    for s, d in zip(net.edge_src.tolist(), net.edge_dst.tolist()):
        adj[s].append(d)
        adj[d].append(s)
    is_ret = (np.bincount(net.edge_dst, minlength=n) >= 2).tolist()
    disc = [-1] * n
    low = [0] * n
    parent = [-1] * n
    edge_stack: List[Tuple[int, int]] = []
    best = 0
    time = 0

    def close(until: Tuple[int, int]) -> None:
        nonlocal best
        nodes = set()
        while edge_stack:
            a, b = edge_stack.pop()
            nodes.add(a)
            nodes.add(b)
            if (a, b) == until or (b, a) == until:
                break
        best = max(best, sum(1 for v in nodes if is_ret[v]))

    for root in range(n):
        if disc[root] != -1:
            continue
        disc[root] = low[root] = time
        time += 1
        stack = [(root, 0)]
        while stack:
            u, k = stack[-1]
            if k < len(adj[u]):
                stack[-1] = (u, k + 1)
                v = adj[u][k]
                if disc[v] == -1:
                    parent[v] = u
                    edge_stack.append((u, v))
                    disc[v] = low[v] = time
                    time += 1
                    stack.append((v, 0))
                elif v != parent[u] and disc[v] < disc[u]:
                    edge_stack.append((u, v))
                    low[u] = min(low[u], disc[v])
            else:
                stack.pop()
                if stack:
                    p = stack[-1][0]
                    low[p] = min(low[p], low[u])
                    if low[u] >= disc[p]:
                        close((p, u))
        if edge_stack:
            close(edge_stack[-1])
    return best
