"""The network / theta prior of MCMC_SEQ on the flat species network.

``log_prior_seq`` of the reference (``src/_mcmc_seq.py:294-411``) is what the informed and coupled reticulation
moves add to every candidate network's score (``_network_only_log_score`` ``:1214``); with the networks flat
(``netmoves.py``) the prior has to be flat too, or every candidate would be rebuilt as a graph just for this sum.
Same terms, same order of accumulation:

* Poisson prior on the number of reticulations, with the ``1 / (E (E - 1))`` topology normaliser per reticulation
  (``E`` = running edge count, starting at ``2 n - 2``),
* Gamma(shape, mean) prior on theta,
* ``Exp(1 / branch_length_mean)`` on every branch length (from node heights, clamped at 0),
* Beta(alpha, beta) on one inheritance probability per reticulation (the first in-edge in edge order; the
  reference takes the first of an unordered edge set -- the same for its symmetric default Beta(2, 2)),
* optionally ``Exp(1 / diameter_mean)`` on every reticulation's diameter |h(parent 1) - h(parent 2)|,
* the identifiability guards: ``-inf`` for a reticulation on an undirected cycle of fewer than ``min_cycle_size``
  nodes or with a hybrid edge shorter than ``min_branch_length``, and beyond ``max_reticulations``.

Host arithmetic only (a few dozen operations per network).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

from ._msnc_density import SpeciesNetwork


@dataclass
class SeqPriors:
    """Hyperparameters, named and defaulted as the reference's ``MCMCSeqPriors`` (``src/_mcmc_seq.py:120-208``)."""

    poisson_mean: float = 1.0
    max_reticulations: int = 4
    max_level: Optional[int] = None
    theta_shape: float = 2.0
    theta_prior_mean: float = 0.036
    theta_mean: float = 0.02
    diameter_mean: float = 0.5
    use_diameter_prior: bool = True
    gamma_alpha: float = 2.0
    gamma_beta: float = 2.0
    min_cycle_size: int = 4
    min_branch_length: float = 1e-6
    branch_length_mean: float = 0.1
    use_branch_length_prior: bool = True


def log_poisson_pmf(k: int, mean: float) -> float:
    if mean <= 0.0:
        return 0.0 if k == 0 else -math.inf
    return k * math.log(mean) - mean - math.lgamma(k + 1)


def log_gamma_pdf(x: float, shape: float, mean: float) -> float:
    if x <= 0.0:
        return -math.inf
    scale = mean / shape
    return (shape - 1.0) * math.log(x) - x / scale - shape * math.log(scale) - math.lgamma(shape)


def log_beta_pdf(x: float, alpha: float, beta: float) -> float:
    if x <= 0.0 or x >= 1.0:
        if alpha == 1.0 and beta == 1.0:
            return 0.0
        return -math.inf
    return ((alpha - 1.0) * math.log(x) + (beta - 1.0) * math.log1p(-x)
            - (math.lgamma(alpha) + math.lgamma(beta) - math.lgamma(alpha + beta)))


def _ancestor_depths(parents_of: List[List[int]], start: int) -> Dict[int, int]:
    dist = {start: 0}
    frontier = [start]
    while frontier:
        nxt = []
        for u in frontier:
            for p in parents_of[u]:
                if p not in dist:
                    dist[p] = dist[u] + 1
                    nxt.append(p)
        frontier = nxt
    return dist


def reticulation_cycle_size(parents_of: List[List[int]], retic: int) -> int:
    """Nodes on the undirected cycle of a reticulation: itself plus the two ancestral paths up to the lowest common
    ancestor of its parents; 2 for parallel edges, 0 for a node that is no proper reticulation (``:264-291``)."""
    parents = parents_of[retic]
    if len(parents) < 2:
        return 0
    p1, p2 = parents[0], parents[1]
    if p1 == p2:
        return 2
    d1 = _ancestor_depths(parents_of, p1)
    d2 = _ancestor_depths(parents_of, p2)
    common = [n for n in d1 if n in d2]
    if not common:
        return 0
    lca = min(common, key=lambda n: d1[n] + d2[n])
    return d1[lca] + d2[lca] + 2


def log_prior_seq(net: SpeciesNetwork, theta: float, priors: SeqPriors) -> float:
    """``log p(Psi) + log p(theta)`` (reference ``log_prior_seq``)."""
    n = net.n_nodes
    src, dst = net.edge_src.tolist(), net.edge_dst.tolist()
    gam = net.edge_gamma.tolist()
    h = net.height.tolist()
    in_edges: List[List[int]] = [[] for _ in range(n)]
    for e, d in enumerate(dst):
        in_edges[d].append(e)
    parents_of = [[src[e] for e in es] for es in in_edges]
    rets = [v for v in range(n) if len(in_edges[v]) >= 2]
    n_ret = len(rets)
    if n_ret > priors.max_reticulations:
        return -math.inf
    for r in rets:
        if priors.min_cycle_size > 0 and reticulation_cycle_size(parents_of, r) < priors.min_cycle_size:
            return -math.inf
        if priors.min_branch_length > 0.0:
            for e in in_edges[r]:
                length = h[src[e]] - h[dst[e]]
                if (length if length > 0.0 else 0.0) < priors.min_branch_length:
                    return -math.inf
    lp = log_poisson_pmf(n_ret, priors.poisson_mean)
    n_leaves = int((np.bincount(net.edge_src, minlength=n) == 0).sum())
    edges = 2 * n_leaves - 2
    for _ in range(n_ret):
        if edges > 1:
            lp -= math.log(edges * (edges - 1))
        edges += 3
    lp += log_gamma_pdf(theta, priors.theta_shape, priors.theta_prior_mean)
    if priors.use_branch_length_prior and priors.branch_length_mean > 0.0:
        rate = 1.0 / priors.branch_length_mean
        log_rate = math.log(rate)
        for s, d in zip(src, dst):
            length = h[s] - h[d]
            if length < 0.0:
                length = 0.0
            lp += log_rate - rate * length
    if n_ret and not (priors.gamma_alpha == 1.0 and priors.gamma_beta == 1.0):
        for r in rets:
            g = gam[in_edges[r][0]]
            if g == g:
                lp += log_beta_pdf(float(g), priors.gamma_alpha, priors.gamma_beta)
    if priors.use_diameter_prior and n_ret:
        rate = 1.0 / priors.diameter_mean
        for r in rets:
            p = parents_of[r]
            lp += math.log(rate) - rate * abs(h[p[0]] - h[p[1]])
    return lp
