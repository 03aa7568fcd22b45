"""Guided joint re-proposal of every locus' gene tree -- the inner machinery of the reference's coupled
add / delete / relocate-reticulation moves (``_coupled_gene_tree_reproposal``, ``src/_mcmc_seq.py:1995-2070``;
``_enumerate_scored_candidates`` ``:1900-1992``) -- with ALL the scoring on the device.

The reference walks the loci one by one; for each it scores ``({g} + NNI(g)) x (0.25, 0.5, 1, 2, 4)`` with a
Felsenstein pruning and an MSNC density per member under the proposed network, samples one member, then
scores the candidate set around the chosen tree under the current network to obtain the reverse density:
``2 * M * K`` prunings and densities on the host per move.  Here one move is FOUR device calls:

1. forward Felsenstein scores of all loci x all members (``score_batch``, ``PNB_EVAL_NOSTORE``),
2. forward MSNC densities of the same trees under the target network (``msnc_batch``),
3. / 4. the same two calls for the candidate sets around the chosen trees, under the reverse network.

The candidate sets themselves are enumerated natively for all loci at once (``candidates.CandidateBatch`` over
``pnb_candidates_count`` / ``pnb_candidates_fill``), directly in the packed form the two calls take.

Sampling follows the reference exactly -- per locus, in locus order, one ``rng.choice(K, p=probs)`` with
``probs = exp(logw - logsumexp(logw))`` renormalised -- and consumes the generator identically, so with the same
seed the same members are chosen whenever the weights agree to rounding.
"""
from __future__ import annotations

import math
from typing import Any, List, Optional, Sequence, Tuple

import numpy as np

from ._msnc_density import LOG_FLOOR
from .candidates import CandidateBatch, _lengths_from_heights, find_member_in_batch
from .tree import FlatTree


def _logsumexp(values: np.ndarray) -> float:
    # reference _logsumexp (src/_msnc_density.py:950-967): floor for an empty / all -inf set, sum in list order
    if values.shape[0] == 0:
        return LOG_FLOOR
    m = float(values.max())
    if m == -math.inf:
        return LOG_FLOOR
    acc = 0.0
    for v in values.tolist():
        acc += math.exp(v - m)
    return m + math.log(acc)


def _member(ft: FlatTree, parent: np.ndarray, blen: np.ndarray) -> FlatTree:
    t = FlatTree.__new__(FlatTree)
    t.parent = np.ascontiguousarray(parent, dtype=np.int32)
    t.blen = np.ascontiguousarray(blen, dtype=np.float64)
    t.label = ft.label
    t._n_children = None
    t._tip_rows = ft._tip_rows       # leaves stay leaves under NNI and scaling
    return t


def _logweights(engine: Any, batch: CandidateBatch, loci: Sequence[int], net: Any, theta: float) -> np.ndarray:
    """``logw[m] = log P(S_i | g_m) + log P(g_m | net, theta)`` for every member of the batch, ``-inf`` where either
    factor is not finite (``_candidate_logweights`` ``:1864-1897``) -- two device calls for all loci."""
    f = engine.score_batch(batch, loci)
    m = engine.msnc_batch(batch, loci, net, theta)
    w = f + m
    w[~(np.isfinite(f) & np.isfinite(m))] = -np.inf
    return w


def coupled_gene_tree_reproposal(engine: Any, gene_trees: Sequence[FlatTree], heights: Sequence[np.ndarray],
                                 target_net: Any, reverse_net: Any, theta: float, rng: np.random.Generator,
                                 choose: Any = None) -> Optional[Tuple[List[FlatTree], List[np.ndarray], float, float]]:
    """Returns ``(new_gene_trees, new_heights, log_qf, log_qr)`` or ``None`` where the reference declines the move
    (a candidate set without a finite-weight member, or an original tree that cannot be matched in the reverse
    set).  ``gene_trees[i]`` / ``heights[i]`` are the current tree of locus ``i`` and its node heights.
    ``choose(i, table, probs) -> j`` replaces the random draw (tests: the member order of a candidate set follows
    node order, which in the reference is the hash order of a set of objects and cannot be replayed)."""
    M = engine.n_loci
    if len(gene_trees) != M or len(heights) != M:
        raise ValueError("coupled_gene_tree_reproposal: one gene tree and one height vector per locus")
    loci = list(range(M))
    ctx = getattr(engine, "_ctx", None)
    # ---- forward: candidate sets of the current trees (one native enumeration), scored under the target network
    fwd = CandidateBatch(gene_trees, heights, context=ctx)
    fwd_lw = _logweights(engine, fwd, loci, target_net, theta)
    new_trees: List[FlatTree] = []
    new_heights: List[np.ndarray] = []
    log_qf = 0.0
    for i in range(M):
        lw = fwd_lw[int(fwd.member_off[i]):int(fwd.member_off[i + 1])]
        lse = _logsumexp(lw)
        if not math.isfinite(lse):
            return None
        probs = np.exp(lw - lse)
        total = float(probs.sum())
        if not math.isfinite(total) or total <= 0.0:
            return None
        probs = probs / total
        j = int(rng.choice(lw.shape[0], p=probs)) if choose is None else int(choose(i, fwd.table(i), probs))
        log_qf += float(lw[j]) - lse
        par, _bl, ht = fwd.member(int(fwd.member_off[i]) + j)
        # materialised as the reference's rebuild(k) does (:1975-1986): lengths re-derived from the heights for
        # EVERY member -- also for the unchanged tree, whose stored lengths may disagree with its heights in the
        # last digits (it was SCORED with its stored lengths, :1828-1832)
        new_trees.append(_member(gene_trees[i], par, _lengths_from_heights(par, ht)))
        new_heights.append(ht.copy())
    # ---- reverse: candidate sets of the chosen trees, scored under the reverse network
    rev = CandidateBatch(new_trees, new_heights, context=ctx)
    rev_lw = _logweights(engine, rev, loci, reverse_net, theta)
    log_qr = 0.0
    for i in range(M):
        lw = rev_lw[int(rev.member_off[i]):int(rev.member_off[i + 1])]
        lse = _logsumexp(lw)
        if not math.isfinite(lse):
            return None
        # the member with the original's clades and (12-decimal) heights: the reference's signature match
        match = find_member_in_batch(rev, i, gene_trees[i].parent, heights[i])
        if match < 0 or not math.isfinite(float(lw[match])):
            return None
        log_qr += float(lw[match]) - lse
    return new_trees, new_heights, log_qf, log_qr


# ---------------------------------------------------------------------------------------------------
# the coupled reticulation moves (network surgery + joint gene-tree re-proposal)
# ---------------------------------------------------------------------------------------------------
def coupled_add_reticulation(engine: Any, net: Any, gene_trees: Sequence[FlatTree], heights: Sequence[np.ndarray],
                             theta: float, rng: np.random.Generator, placements_for: Any, *, placement: Any = None,
                             choose: Any = None, max_level: Optional[int] = None
                             ) -> Optional[Tuple[Any, List[FlatTree], List[np.ndarray], float]]:
    """``op_add_reticulation_coupled`` (``src/_mcmc_seq.py:2072-2192``) on flat objects: the informed add move of
    ``netmoves`` followed by the joint gene-tree re-proposal under the NEW network (reverse density under the old
    one).  ``placements_for(network, gene_trees)`` returns the scored placements
    (``netmoves.add_reticulation_placements`` with the caller's prior).  Returns
    ``(new_network, new_gene_trees, new_heights, log HR)`` with

        log HR = [informed-add ratio] + (log q_r - log q_f)                                   (``:2168-2175``)

    ``placement=(k, u_t2, u_t1, gamma)`` / ``choose`` replace the random draws (tests)."""
    from . import netmoves as NM
    placements = placements_for(net, gene_trees)
    if not placements:
        return None
    mv = NM.add_reticulation_move(net, placements, rng) if placement is None else NM.add_reticulation_at(net, placements, *placement)
    if mv is None:
        return None
    new_net, loghr_net = mv
    if max_level is not None and NM.network_level(new_net) > max_level:      # before the expensive part (:2150-2155)
        return None
    repro = coupled_gene_tree_reproposal(engine, gene_trees, heights, new_net, net, theta, rng, choose)
    if repro is None:
        return None
    new_trees, new_heights, log_qf, log_qr = repro
    return new_net, new_trees, new_heights, loghr_net + (log_qr - log_qf)


def coupled_delete_reticulation(engine: Any, net: Any, gene_trees: Sequence[FlatTree], heights: Sequence[np.ndarray],
                                theta: float, rng: np.random.Generator, placements_for: Any, *, pick: Any = None,
                                choose: Any = None) -> Optional[Tuple[Any, List[FlatTree], List[np.ndarray], float]]:
    """``op_delete_reticulation_coupled`` (``:2195-2330``): remove a reticulation (``pick=(retic, donor_parent)`` or
    drawn as the reference draws), re-propose every gene tree under the post-delete network, then score the reverse
    add's placements on the post-delete network with the NEW gene trees (``:2293-2299``)."""
    from . import netmoves as NM
    if pick is None:
        pick = NM.draw_reticulation_to_delete(net, rng)
        if pick is None:
            return None
    geom = NM.delete_reticulation_geometry(net, int(pick[0]), int(pick[1]))
    if geom is None:
        return None
    new_net = geom["network"]
    repro = coupled_gene_tree_reproposal(engine, gene_trees, heights, new_net, net, theta, rng, choose)
    if repro is None:
        return None
    new_trees, new_heights, log_qf, log_qr = repro
    loghr_net = NM.delete_reticulation_loghr(geom, placements_for(new_net, new_trees))
    if loghr_net is None:
        return None
    return new_net, new_trees, new_heights, loghr_net + (log_qr - log_qf)


def coupled_relocate_reticulation(engine: Any, net: Any, gene_trees: Sequence[FlatTree], heights: Sequence[np.ndarray],
                                  theta: float, rng: np.random.Generator, *, pick: Any = None, placement: Any = None,
                                  choose: Any = None, max_level: Optional[int] = None
                                  ) -> Optional[Tuple[Any, List[FlatTree], List[np.ndarray], float]]:
    """``op_relocate_reticulation_coupled`` (``src/_mcmc_seq.py:2604-2790``): detach a reticulation A (as the delete
    move does), re-attach it at a placement B drawn UNIFORMLY from the geometric placements of the base network,
    re-propose every gene tree under the new network.  The ``1 / K`` selection cancels between the two directions,
    so (``_relocation_loghr`` ``:2578-2601``)

        log HR = log l2_B + log win_B - log l2_A - log win_A + (log q_r - log q_f).

    ``pick=(retic, donor_parent)``, ``placement=(donor_edge, retic_edge, u_t2, u_t1, gamma)`` (edges of the base
    network) and ``choose`` replace the random draws (tests); ``max_level`` is the level cap of ``MCMCSeqPriors``."""
    from . import netmoves as NM
    if pick is None:
        pick = NM.draw_reticulation_to_delete(net, rng)
        if pick is None:
            return None
    geom = NM.delete_reticulation_geometry(net, int(pick[0]), int(pick[1]))
    if geom is None:
        return None
    n0 = geom["network"]
    if placement is None:
        places = NM.geometric_placements(n0)
        if not places:
            return None
        d, r = places[int(rng.integers(len(places)))]
        u2 = u1 = g = None
    else:
        d, r, u2, u1, g = placement
    h = n0.height
    b3, b4 = int(n0.edge_src[d]), int(n0.edge_dst[d])
    b5, b6 = int(n0.edge_src[r]), int(n0.edge_dst[r])
    l2_b = float(h[b5] - h[b6])
    if l2_b <= 0.0:
        return None
    if u2 is None:
        u2 = float(rng.random())
    t2 = float(h[b6]) + l2_b * u2
    lo = max(float(h[b4]), t2)
    win_b = float(h[b3]) - lo
    if win_b <= 0.0:
        return None
    if u1 is None:
        u1 = float(rng.random())
        g = float(rng.random())
    new_net = NM.insert_reticulation(n0, int(d), int(r), lo + win_b * u1, t2, float(g))
    if new_net is None:
        return None
    if max_level is not None and NM.network_level(new_net) > max_level:
        return None
    repro = coupled_gene_tree_reproposal(engine, gene_trees, heights, new_net, net, theta, rng, choose)
    if repro is None:
        return None
    new_trees, new_heights, log_qf, log_qr = repro
    loghr = (math.log(l2_b) + math.log(win_b) - math.log(geom["l2"]) - math.log(geom["win"]) + (log_qr - log_qf))
    return new_net, new_trees, new_heights, loghr
