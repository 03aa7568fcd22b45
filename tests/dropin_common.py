"""Shared pieces of the drop-in tests (CPU: fake device; GPU: the kernels): the UNMODIFIED reference
(``baseline/_ref``) runs its own ``MCMC_SEQ`` on top of ``phynetpy_b200.dropin``.

The reference's chain is not reproducible from its seed (set iteration over ``Edge`` objects hashes
by address: two runs of the unpatched reference with one seed already differ), so "same accept /
reject sequence" cannot be tested by comparing two runs.  It is tested inside ONE run instead:
:class:`LockstepEngine` evaluates every likelihood the chain asks for with BOTH implementations --
the reference's own ``_SeqLikelihoodEngine`` / ``FelsensteinCalculator`` and the drop-in -- on the
same proposal, and the chain loop (a restatement of the loop body of ``MCMC_SEQ.search``,
``src/_mcmc_seq.py:3478-3499``, over the reference's own ``SeqState`` and ``MCMCSeqKernel``) takes
the accept / reject decision with both values and the same uniform draw.
"""
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from baseline import ref_loader  # noqa: E402


def require_reference():
    if not ref_loader.available():
        pytest.skip("baseline/_ref is not installed (baseline/install_reference.sh)")
    return ref_loader.load_reference()


def make_loci(seed=0, n_sites=120, n_loci=3, taxa="ABCD"):
    """The data recipe of the reference's own driver tests (``tests/test_mcmc_seq.py:289-308, 358-381``)."""
    rng = np.random.default_rng(seed)

    def rand(n):
        return "".join(rng.choice(list("ACGT"), size=n))

    def mutate(s, k):
        s = list(s)
        for _ in range(k):
            s[rng.integers(len(s))] = rng.choice(list("ACGT"))
        return "".join(s)

    def locus():
        core = rand(n_sites)
        seqs = {}
        prev = core
        for k, t in enumerate(taxa):
            prev = mutate(core if k >= 2 else prev, 5 + 15 * max(k - 1, 0))
            seqs[t] = prev
        return seqs

    return [locus() for _ in range(n_loci)], {t: [t] for t in taxa}


def make_lockstep_engine_class(mcmc_mod, ref_engine_cls, ref_calc_cls, log):
    """A ``B200SeqLikelihoodEngine`` that shadows every call with the reference's own engine."""
    from phynetpy_b200 import dropin

    class _LockstepCalc:
        def __init__(self, b200_calc, ref_calc):
            self._b, self._r = b200_calc, ref_calc

        def __getattr__(self, name):
            return getattr(self._b, name)

        def log_likelihood(self, gene_tree, model, **kw):
            b = self._b.log_likelihood(gene_tree, model, **kw)
            r = self._r.log_likelihood(gene_tree, model, **kw)
            log["candidates"].append((b, r))
            return b

    class LockstepEngine(dropin.B200SeqLikelihoodEngine):
        def __init__(self, loci, species_of, model, branch_thetas=None):
            super().__init__(loci, species_of, model, branch_thetas)
            saved = mcmc_mod.FelsensteinCalculator
            mcmc_mod.FelsensteinCalculator = ref_calc_cls      # the shadow is the reference, untouched
            try:
                self._shadow = ref_engine_cls(loci, species_of, model, branch_thetas)
            finally:
                mcmc_mod.FelsensteinCalculator = saved
            self._calcs = [_LockstepCalc(c, rc) for c, rc in zip(self._calcs, self._shadow._calcs)]
            self.ref_total = None

        def invalidate_locus(self, i):
            super().invalidate_locus(i)
            self._shadow.invalidate_locus(i)

        def invalidate_network(self):
            super().invalidate_network()
            self._shadow.invalidate_network()

        def clone_caches(self):
            snap = super().clone_caches()
            return (snap[0], snap[1], self._shadow.clone_caches())

        def restore_caches(self, snapshot):
            super().restore_caches(snapshot)
            self._shadow.restore_caches(snapshot[2])

        def log_likelihood(self, gene_trees, species_net, theta):
            try:
                b = super().log_likelihood(gene_trees, species_net, theta)
            except Exception as exc:   # MCMCSeqKernel.propose would swallow it into a reject (:2878-2881)
                log["errors"].append(repr(exc))
                raise
            r = self._shadow.log_likelihood(gene_trees, species_net, theta)
            self.ref_total = r
            log["evals"].append((b, r))
            for x, y in zip(self._felsen, self._shadow._felsen):
                log["felsen"].append((x, y))
            for x, y in zip(self._msnc, self._shadow._msnc):
                log["msnc"].append((x, y))
            return b

    return LockstepEngine


def close(a, b, rel=1e-10):
    if a == b:
        return True
    if not (math.isfinite(a) and math.isfinite(b)):
        return False
    return abs(a - b) <= rel * max(1.0, abs(b))


def run_lockstep_chain(mcmc_mod, loci, mapping, n_iter, seed, max_reticulations=2, coupled=True):
    """The loop body of ``MCMC_SEQ.search`` (``src/_mcmc_seq.py:3478-3499``) with both likelihood values."""
    from phynetpy_b200 import dropin
    assert dropin.installed()
    log = {"evals": [], "candidates": [], "felsen": [], "msnc": [], "errors": [], "decisions": [], "placements": []}
    undo = dropin._state["undo"]
    ref_engine_cls = [v for (m, a, v) in undo if a == "_SeqLikelihoodEngine"][0]
    ref_calc_cls = [v for (m, a, v) in undo if a == "FelsensteinCalculator" and m is mcmc_mod][0]
    saved = mcmc_mod._SeqLikelihoodEngine
    mcmc_mod._SeqLikelihoodEngine = make_lockstep_engine_class(mcmc_mod, ref_engine_cls, ref_calc_cls, log)
    # the placement scan of the informed add / delete / relocate moves (_network_only_log_score, :1163-1214): every
    # candidate network scored by the drop-in (one device call) AND by the reference's own host loop
    saved_scan = mcmc_mod._network_only_log_score
    ref_scan = dropin._state["ref_network_only_log_score"]

    def both_scans(state, net):
        b = saved_scan(state, net)
        log["placements"].append((b, ref_scan(state, net)))
        return b
    mcmc_mod._network_only_log_score = both_scans
    try:
        inf = mcmc_mod.MCMC_SEQ(loci, mapping, priors=mcmc_mod.MCMCSeqPriors(max_reticulations=max_reticulations))
        rng = np.random.default_rng(seed)
        kernel = mcmc_mod.MCMCSeqKernel(rng, max_reticulations=max_reticulations, coupled_dimension_moves=coupled)
        state = inf._new_state()
        eng = state._engine
        cur_b = state.log_posterior()
        cur_r = eng.ref_total + state.log_prior()
        accepted = 0
        for it in range(n_iter):
            # MCMCSeqKernel.propose (:2867-2881) with its exception swallow made visible: an exception that
            # comes out of the drop-in (not out of the reference's own operator logic) is a test failure
            idx = int(kernel.rng.choice(len(kernel._ops), p=kernel._weights))
            try:
                proposal = kernel._ops[idx][0](state, kernel.rng)
            except Exception as exc:
                import traceback
                tb = traceback.format_exc()
                if "phynetpy_b200" in tb or "B200" in repr(exc) or "Lockstep" in repr(exc):
                    log["errors"].append(tb)
                proposal = None
            if proposal is None:
                continue
            loghr, undo_fn = proposal
            prop_b = state.log_posterior()
            prop_r = eng.ref_total + state.log_prior()
            if not math.isfinite(prop_b) or not math.isfinite(prop_r):
                log["decisions"].append((math.isfinite(prop_b), math.isfinite(prop_r)))
                undo_fn()
                continue
            u = math.log(rng.random())
            acc_b = u < (prop_b - cur_b) + loghr
            acc_r = u < (prop_r - cur_r) + loghr
            log["decisions"].append((acc_b, acc_r))
            if acc_b:
                cur_b, cur_r = prop_b, prop_r
                accepted += 1
            else:
                undo_fn()
        # what is cached after the chain equals a from-scratch evaluation of the final state
        final_cached = state.log_likelihood()
        for i in range(eng.n_loci):
            eng.invalidate_locus(i)
        eng.invalidate_network()
        eng.invalidate_device_cache()
        final_fresh = state.log_likelihood()
        log["final"] = (final_cached, final_fresh, eng.ref_total)
        log["accepted"] = accepted
        log["reticulations_seen"] = state.num_reticulations()
        return log
    finally:
        mcmc_mod._SeqLikelihoodEngine = saved
        mcmc_mod._network_only_log_score = saved_scan


def check_lockstep_log(log, n_iter):
    assert not log["errors"], log["errors"][:3]
    assert len(log["evals"]) > n_iter // 2
    for b, r in log["evals"]:
        assert close(b, r), (b, r)
    for b, r in log["candidates"]:
        assert close(b, r), (b, r)
    for b, r in log["felsen"]:
        assert close(b, r), (b, r)
    for b, r in log["msnc"]:
        assert close(b, r), (b, r)
    for b, r in log["placements"]:
        assert close(b, r), (b, r)
    assert len(log["placements"]) > 20, "the placement scan never ran through the drop-in"
    assert all(a == b for a, b in log["decisions"]), "an accept / reject decision differs"
    assert 0 < log["accepted"] < len(log["decisions"])
    cached, fresh, ref = log["final"]
    assert close(cached, fresh, 1e-12) and close(fresh, ref)
