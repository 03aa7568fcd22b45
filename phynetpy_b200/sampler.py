"""A single-chain MCMC_SEQ driver on flat objects -- the thin loop that strings the validated pieces together.

The reference's sampler (``MCMC_SEQ`` / ``MCMCSeqKernel`` / ``SeqState``, ``src/_mcmc_seq.py:775-880,2800-2880,
3175-3580``) clones and mutates ``Network`` graphs on every proposal; this one keeps the whole state flat
(:class:`~phynetpy_b200.tree.FlatTree` gene trees + node heights, a flat
:class:`~phynetpy_b200._msnc_density.SpeciesNetwork`, theta) and calls

* ``SeqLikelihoodEngine.log_likelihood(gene_trees, species_net, theta)`` for both likelihood factors (device),
* ``moves`` / ``netmoves`` / ``coupled`` for the proposals and their log Hastings ratios,
* ``priors.log_prior_seq`` for the prior,
* ``start`` for the default starting state (UPGMA gene trees, UPGMA species tree on the per-species
  concatenation, embedding-consistency shrink).

Same operator mixture and weights as ``MCMCSeqKernel`` (gene node height 0.28, gene NNI 0.18, network node height
0.18, gamma 0.07, theta 0.09, add 0.06, delete 0.06, relocate 0.08; add / delete are the network-only informed moves
unless ``coupled_dimension_moves``), the same accept rule (``log u < (post' - post) + log HR``; a proposal that
cannot be built, or whose posterior is not finite, is a rejection) and the same copy-and-swap discipline: a
rejected proposal is undone by ``restore_caches``, which also rolls the pending device partials back.

Deliberately small: a single chain (:meth:`FlatMCMCSeq.search`) and the MC3 ladder in lockstep
(:meth:`FlatMCMCSeq.search_mc3`, all chains' dirty loci in one device call); no warm start from MCMC_GT, no control
hooks, no trace files -- those stay the reference's.  Random
streams are not comparable with the reference's (its node enumeration follows the hash order of Python sets).
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import coupled as CP
from . import moves as MV
from . import netmoves as NM
from . import start as ST
from ._engine import SeqLikelihoodEngine
from ._msnc_density import SpeciesNetwork
from .candidates import HEIGHT_SCALE_GRID, node_heights
from .priors import SeqPriors, log_prior_seq
from .tree import FlatTree

OPERATORS = (("gene_node_height", 0.28), ("gene_tree_nni", 0.18), ("net_node_height", 0.18), ("change_gamma", 0.07),
             ("change_theta", 0.09), ("add_reticulation", 0.06), ("delete_reticulation", 0.06),
             ("relocate_reticulation", 0.08))


@dataclass
class FlatSample:
    iteration: int
    log_posterior: float
    log_likelihood: float
    theta: float
    num_reticulations: int
    gamma_major: Optional[float]
    network: SpeciesNetwork


@dataclass
class FlatResult:
    map_network: SpeciesNetwork
    map_log_posterior: float
    map_theta: float
    map_gene_trees: List[FlatTree]
    samples: List[FlatSample] = field(default_factory=list)
    acceptance_rate: float = 0.0
    num_iterations: int = 0
    wall_time_sec: float = 0.0
    operator_counts: Dict[str, List[int]] = field(default_factory=dict)    # name -> [proposed, accepted]


def species_tree_from_concatenation(loci: Sequence[Dict[str, str]], mapping: Dict[str, List[str]]) -> SpeciesNetwork:
    """``_build_species_tree`` (``src/_mcmc_seq.py:3147-3172``): UPGMA on the concatenation of all loci, one
    representative allele per species."""
    concat: Dict[str, str] = {}
    for sp, alleles in mapping.items():
        pieces = []
        for aln in loci:
            seq = aln.get(alleles[0])
            if seq is None:
                seq = next((aln[a] for a in alleles if a in aln), None)
            pieces.append(seq if seq is not None else "")
        concat[sp] = "".join(pieces)
    ft = ST.build_upgma_gene_tree(concat)
    h = node_heights(ft)
    nz = ft.parent >= 0
    labels = [lab if lab is not None else "I%d" % i for i, lab in enumerate(ft.label)]
    leaf = [lab for lab in ft.label]
    return SpeciesNetwork(ft.parent[nz].tolist(), np.nonzero(nz)[0].tolist(), [None] * int(nz.sum()), h, leaf, labels)


class FlatMCMCSeq:
    """``MCMC_SEQ(loci, mapping, priors=..., model=...)`` with a flat state; ``engine`` may be supplied (the CPU tests pass
    one over a fake device), otherwise a :class:`SeqLikelihoodEngine` on the default device is built."""

    def __init__(self, loci: Sequence[Dict[str, str]], mapping: Dict[str, List[str]], *, priors: Optional[SeqPriors] = None,
                 model: Any = None, theta: Optional[float] = None, gene_trees: Optional[Sequence[FlatTree]] = None,
                 species_net: Optional[SpeciesNetwork] = None, engine: Any = None, coupled_dimension_moves: bool = False):
        from ._seq_likelihood import JC69
        self.loci = list(loci)
        self.mapping = {sp: list(al) for sp, al in mapping.items()}
        self.priors = priors if priors is not None else SeqPriors()
        self.model = model if model is not None else JC69()
        self.theta = float(theta if theta is not None else self.priors.theta_mean)
        if not (math.isfinite(self.theta) and self.theta > 0.0):
            raise ValueError("theta must be finite and positive; got %s." % self.theta)
        self.species_of = {a: sp for sp, al in self.mapping.items() for a in al}
        self.trees = list(gene_trees) if gene_trees is not None else [ST.build_upgma_gene_tree(a) for a in self.loci]
        self.heights = [node_heights(t) for t in self.trees]
        net = species_net if species_net is not None else species_tree_from_concatenation(self.loci, self.mapping)
        # independently built starting trees rarely embed: shrink the network until they do (SeqState.__init__)
        self.net, _ = ST.enforce_embedding_consistency(net, self.trees, self.heights, self.species_of)
        self.engine = engine if engine is not None else SeqLikelihoodEngine(self.loci, self.species_of, self.model)
        self.coupled_dimension_moves = coupled_dimension_moves

    # -- posterior -----------------------------------------------------------------------------------
    def log_likelihood(self) -> float:
        return self.engine.log_likelihood(self.trees, self.net, self.theta)

    def log_prior(self) -> float:
        return log_prior_seq(self.net, self.theta, self.priors)

    def score(self) -> float:
        return self.log_likelihood() + self.log_prior()

    def _placements_for(self, net: SpeciesNetwork, trees: Sequence[FlatTree]):
        dens = NM.device_density_fn(trees, self.species_of, self.theta, context=self.engine._context())
        pri, theta = self.priors, self.theta
        return NM.add_reticulation_placements(net, dens, len(HEIGHT_SCALE_GRID), prior_fn=lambda n: log_prior_seq(n, theta, pri))

    # -- one proposal ----------------------------------------------------------------------------------
    def _propose(self, name: str, rng: np.random.Generator):
        """Returns ``None`` or ``(log HR, net', trees', heights', theta', dirty loci | 'all' | None, network dirty)``."""
        eng = self.engine
        M = len(self.trees)
        if name in ("gene_node_height", "gene_tree_nni"):
            if M == 0:
                return None
            i = int(rng.integers(M))
            if name == "gene_node_height":
                mv = MV.slide_height_move(self.trees[i], self.heights[i], rng)
            else:
                nn = MV.nni_move(self.trees[i], self.heights[i], rng)
                mv = None if nn is None else (nn[0], nn[1], 0.0)
            if mv is None:
                return None
            trees, heights = list(self.trees), list(self.heights)
            trees[i], heights[i] = mv[0], mv[1]
            return mv[2], self.net, trees, heights, self.theta, [i], False
        if name == "net_node_height":
            mv = MV.net_slide_height_move(self.net, rng)
            return None if mv is None else (mv[1], mv[0], self.trees, self.heights, self.theta, None, True)
        if name == "change_gamma":
            mv = MV.gamma_move(self.net, rng)
            return None if mv is None else (mv[1], mv[0], self.trees, self.heights, self.theta, None, True)
        if name == "change_theta":
            th, lh = MV.theta_move(self.theta, rng)
            return lh, self.net, self.trees, self.heights, th, None, True
        R = self.net.n_reticulations()
        lvl = self.priors.max_level
        if name == "add_reticulation":
            if R >= self.priors.max_reticulations:
                return None
            if self.coupled_dimension_moves:
                out = CP.coupled_add_reticulation(eng, self.net, self.trees, self.heights, self.theta, rng, self._placements_for,
                                                  max_level=lvl)
                return None if out is None else (out[3], out[0], out[1], out[2], self.theta, "all", True)
            mv = NM.add_reticulation_move(self.net, self._placements_for(self.net, self.trees), rng)
            if mv is None or (lvl is not None and NM.network_level(mv[0]) > lvl):
                return None
            return mv[1], mv[0], self.trees, self.heights, self.theta, None, True
        if name == "delete_reticulation":
            if R == 0:
                return None
            if self.coupled_dimension_moves:
                out = CP.coupled_delete_reticulation(eng, self.net, self.trees, self.heights, self.theta, rng, self._placements_for)
                return None if out is None else (out[3], out[0], out[1], out[2], self.theta, "all", True)
            mv = NM.delete_reticulation_move(self.net, rng, lambda n: self._placements_for(n, self.trees))
            return None if mv is None else (mv[1], mv[0], self.trees, self.heights, self.theta, None, True)
        if name == "relocate_reticulation":
            if R == 0:
                return None
            out = CP.coupled_relocate_reticulation(eng, self.net, self.trees, self.heights, self.theta, rng, max_level=lvl)
            return None if out is None else (out[3], out[0], out[1], out[2], self.theta, "all", True)
        raise ValueError("unknown operator %r" % name)

    # -- the chain --------------------------------------------------------------------------------------
    def _apply(self, prop) -> None:
        """Tell the engine what a proposal changed (its evaluation then recomputes exactly that)."""
        _loghr, _net, trees, _heights, _theta, dirty, net_dirty = prop
        eng = self.engine
        if dirty == "all":
            for i in range(len(trees)):
                eng.invalidate_locus(i)
        elif dirty:
            for i in dirty:
                eng.invalidate_locus(i)
        if net_dirty:
            eng.invalidate_network()

    def _adopt(self, prop, new: float) -> None:
        _loghr, net, trees, heights, theta, _dirty, _net_dirty = prop
        self.net, self.trees, self.heights, self.theta = net, list(trees), list(heights), theta

    def _sample(self, it: int, cur: float) -> FlatSample:
        rets = MV.reticulations(self.net).tolist()
        gmaj = None
        if rets:
            g = self.net.edge_gamma[np.isin(self.net.edge_dst, rets)]
            g = g[g == g]
            gmaj = float(g.max()) if g.shape[0] else None
        return FlatSample(it, cur, cur - self.log_prior(), self.theta, len(rets), gmaj, self.net)

    def search(self, *, num_iter: int = 20000, burn_in: int = 5000, sample_freq: int = 100, seed: Any = None) -> FlatResult:
        rng = np.random.default_rng(seed)
        names = [n for n, _ in OPERATORS]
        weights = np.asarray([w for _, w in OPERATORS], dtype=np.float64)
        weights /= weights.sum()
        eng = self.engine
        t0 = time.perf_counter()
        cur = self.score()
        if not math.isfinite(cur):
            raise RuntimeError("the starting state has zero posterior (log posterior %r)" % cur)
        best = (cur, self.net, self.theta, list(self.trees))
        counts = {n: [0, 0] for n in names}
        samples: List[FlatSample] = []
        accepted = 0
        for it in range(num_iter):   # 0-based, as the reference (src/_mcmc_seq.py:3478, :3637)
            name = names[int(rng.choice(len(names), p=weights))]
            counts[name][0] += 1
            snap = eng.clone_caches()
            try:
                prop = self._propose(name, rng)
            except Exception:
                prop = None                          # an operator that fails is a rejection (:2878-2881)
                eng.restore_caches(snap)
            if prop is not None:
                self._apply(prop)
                new = eng.log_likelihood(prop[2], prop[1], prop[4]) + log_prior_seq(prop[1], prop[4], self.priors)
                if math.isfinite(new) and math.log(rng.random()) < (new - cur) + prop[0]:
                    self._adopt(prop, new)
                    cur = new
                    accepted += 1
                    counts[name][1] += 1
                    if cur > best[0]:
                        best = (cur, self.net, self.theta, list(self.trees))
                else:
                    eng.restore_caches(snap)
            if it >= burn_in and sample_freq > 0 and (it - burn_in) % sample_freq == 0:   # (:3501, :3668)
                samples.append(self._sample(it, cur))
        return FlatResult(best[1], best[0], best[2], best[3], samples, accepted / max(num_iter, 1), num_iter,
                          time.perf_counter() - t0, counts)

    # -- Metropolis-coupled chains (MC3) ---------------------------------------------------------------------
    def fork(self, engine: Any) -> "FlatMCMCSeq":
        """A second chain on the same data and starting state, with its own engine (its own partials cache)."""
        other = FlatMCMCSeq.__new__(FlatMCMCSeq)
        other.__dict__.update(self.__dict__)
        other.trees, other.heights = list(self.trees), list(self.heights)
        other.engine = engine
        return other

    def search_mc3(self, temperatures: Sequence[float], *, engines: Optional[Sequence[Any]] = None, num_iter: int = 20000,
                   burn_in: int = 5000, sample_freq: int = 100, swap_interval: int = 1, seed: Any = None) -> FlatResult:
        """The reference's MC3 ladder (``_mc3``, ``src/_mcmc_seq.py:3584-3723``) in lockstep: chain c targets
        ``posterior ** (1 / T_c)`` (accept if ``log u < (post' - post) / T_c + log HR``), adjacent chains swap states
        with log ratio ``(1/T_i - 1/T_j) (post_j - post_i)`` every ``swap_interval`` iterations, the coldest chain is
        sampled.  What differs is the evaluation: all chains propose first, then the dirty loci of ALL chains are
        scored in one device call (``SeqLikelihoodEngine.log_likelihood_many``) -- T chains, one launch.
        ``engines``: one engine per additional chain (built on the default device when omitted)."""
        m = len(temperatures)
        if m < 2:
            return self.search(num_iter=num_iter, burn_in=burn_in, sample_freq=sample_freq, seed=seed)
        betas = [1.0 / float(t) for t in temperatures]
        cold = int(min(range(m), key=lambda i: temperatures[i]))
        if engines is None:
            engines = [SeqLikelihoodEngine(self.loci, self.species_of, self.model) for _ in range(m - 1)]
        if len(engines) != m - 1:
            raise ValueError("search_mc3: one engine per additional chain")
        chains = [self] + [self.fork(e) for e in engines]
        ss = seed if isinstance(seed, np.random.SeedSequence) else np.random.SeedSequence(seed)
        rngs = [np.random.default_rng(s) for s in ss.spawn(m)]
        swap_rng = np.random.default_rng(ss.spawn(1)[0])
        names = [n for n, _ in OPERATORS]
        weights = np.asarray([w for _, w in OPERATORS], dtype=np.float64)
        weights /= weights.sum()
        many = type(self.engine).log_likelihood_many
        t0 = time.perf_counter()
        liks = many([c.engine for c in chains], [c.trees for c in chains], [c.net for c in chains], [c.theta for c in chains])
        curs = [lk + c.log_prior() for lk, c in zip(liks, chains)]
        if not all(math.isfinite(x) for x in curs):
            raise RuntimeError("the starting state has zero posterior")
        best = (curs[cold], chains[cold].net, chains[cold].theta, list(chains[cold].trees))
        samples: List[FlatSample] = []
        counts = {n: [0, 0] for n in names}
        accepted = swaps = swaps_ok = 0
        for it in range(num_iter):   # 0-based, as the reference (src/_mcmc_seq.py:3478, :3637)
            props: List[Any] = []
            snaps = []
            for c, ch in enumerate(chains):
                name = names[int(rngs[c].choice(len(names), p=weights))]
                if c == cold:
                    counts[name][0] += 1
                snaps.append(ch.engine.clone_caches())
                try:
                    prop = ch._propose(name, rngs[c])
                except Exception:
                    prop = None
                    ch.engine.restore_caches(snaps[-1])
                if prop is not None:
                    ch._apply(prop)
                props.append((name, prop))
            # one device call for the dirty loci of every chain (a chain without a proposal has nothing dirty)
            liks = many([ch.engine for ch in chains],
                        [p[1][2] if p[1] is not None else ch.trees for p, ch in zip(props, chains)],
                        [p[1][1] if p[1] is not None else ch.net for p, ch in zip(props, chains)],
                        [p[1][4] if p[1] is not None else ch.theta for p, ch in zip(props, chains)])
            for c, ch in enumerate(chains):
                name, prop = props[c]
                if prop is None:
                    continue
                new = liks[c] + log_prior_seq(prop[1], prop[4], ch.priors)
                if math.isfinite(new) and math.log(rngs[c].random()) < betas[c] * (new - curs[c]) + prop[0]:
                    ch._adopt(prop, new)
                    curs[c] = new
                    if c == cold:
                        accepted += 1
                        counts[name][1] += 1
                else:
                    ch.engine.restore_caches(snaps[c])
            if curs[cold] > best[0]:
                best = (curs[cold], chains[cold].net, chains[cold].theta, list(chains[cold].trees))
            if it % swap_interval == 0:
                i = int(swap_rng.integers(0, m - 1))
                j = i + 1
                swaps += 1
                if math.log(swap_rng.random()) < (betas[i] - betas[j]) * (curs[j] - curs[i]):
                    # the states migrate, each temperature slot keeps its generator; a state travels WITH its engine,
                    # whose device caches describe exactly that state
                    chains[i], chains[j] = chains[j], chains[i]
                    curs[i], curs[j] = curs[j], curs[i]
                    swaps_ok += 1
            if it >= burn_in and sample_freq > 0 and (it - burn_in) % sample_freq == 0:   # (:3501, :3668)
                samples.append(chains[cold]._sample(it, curs[cold]))
        # leave this object in the cold chain's final state
        final = chains[cold]
        self.net, self.trees, self.heights, self.theta, self.engine = final.net, final.trees, final.heights, final.theta, final.engine
        res = FlatResult(best[1], best[0], best[2], best[3], samples, accepted / max(num_iter, 1), num_iter,
                         time.perf_counter() - t0, counts)
        res.operator_counts["__swaps__"] = [swaps, swaps_ok]
        return res
