"""CPU: the reference's own MCMC_SEQ (unmodified, ``baseline/_ref``) on top of ``phynetpy_b200.dropin`` with a
fake device that answers with the oracles and enforces the C ABI's state rules.  What is tested here is the host
side of the drop-in under the REAL driver: the patch itself, the pending / commit / rollback protocol under every
operator and undo closure of the reference, the score-only route of the coupled moves' direct Felsenstein calls.
The same scenarios run against the kernels in ``tests/test_gpu_dropin.py``."""
import math
import types

import numpy as np
import pytest

from oracle import felsenstein_oracle as O, msnc_oracle as MO
from phynetpy_b200 import _engine as E, _lib, _seq_likelihood as SL, dropin
import dropin_common as DC  # tests/ is on sys.path (rootdir conftest)


class FakeDevice:
    """Stands in for DeviceContext; loci register themselves (handle = index)."""

    def __init__(self, model):
        self.loci = []
        self.model = model
        self.pending = set()
        self.calls = []
        self._h = object()

    def register(self, calc):
        self.loci.append(O.OracleCodesLocus(calc._codes, calc._int_weights))
        return len(self.loci) - 1

    def eval(self, model_handle, handles, node_off, parent, blen, tip_row, site_rate=1.0, flags=0, P_edge=None):
        out = np.empty(len(handles))
        nostore = bool(flags & _lib.EVAL_NOSTORE)
        seen = set()
        for j, h in enumerate(handles.tolist()):
            if not nostore:
                assert h not in self.pending, "locus %d evaluated while an evaluation is pending (PNB_ERR_STATE)" % h
                assert h not in seen, "a locus may appear twice in a batch only with NOSTORE"
                seen.add(h)
                assert flags & _lib.EVAL_PENDING, "the engine evaluates proposals PENDING"
                self.pending.add(h)
            a, b = int(node_off[j]), int(node_off[j + 1])
            tree = O.OracleTree(parent[a:b].tolist(), [None if x != x else x for x in blen[a:b].tolist()], [None] * (b - a))
            out[j] = self.loci[h].log_likelihood(tip_row[a:b].tolist(), tree, self.model, site_rate)
        self.calls.append(("eval", len(handles), int(flags)))
        return out

    def commit(self, handles):
        for h in handles.tolist():
            assert h in self.pending, "commit of a locus without a pending evaluation"
            self.pending.discard(h)
        self.calls.append(("commit", len(handles)))

    def rollback(self, handles):
        for h in handles.tolist():
            assert h in self.pending, "rollback of a locus without a pending evaluation"
            self.pending.discard(h)
        self.calls.append(("rollback", len(handles)))

    def last_stats(self):
        return {"launches": 1, "updates": 0, "nodes_recomputed": 0}


@pytest.fixture
def world(monkeypatch):
    ref = DC.require_reference()
    import phynetpy._mcmc_seq as mcmc
    dev = FakeDevice(O.jc69())
    monkeypatch.setattr(E, "default_context", lambda device=None: dev)
    monkeypatch.setattr(E, "_model_handle", lambda model, ctx: None)
    monkeypatch.setattr(SL.FelsensteinCalculator, "_locus_handle",
                        lambda self: self.__dict__.setdefault("_fake", types.SimpleNamespace(value=dev.register(self))))

    def fake_msnc(net, off, parent, blen, ids, theta, ctx=None, branch_thetas=None):
        netd = net.to_dict()
        out = np.empty(len(off) - 1)
        for j in range(len(off) - 1):
            a, b = int(off[j]), int(off[j + 1])
            labels = [net.species[s] if s >= 0 else None for s in ids[a:b].tolist()]
            out[j] = MO.msnc_log_density(netd, parent[a:b].tolist(), blen[a:b].tolist(), labels,
                                         {sp: sp for sp in net.species}, theta, branch_thetas)
        return out
    monkeypatch.setattr(E, "msnc_eval_packed", fake_msnc)
    dropin.install(ref)
    yield types.SimpleNamespace(ref=ref, mcmc=mcmc, dev=dev)
    dropin.uninstall()


def test_install_replaces_the_seam_and_uninstall_restores_it():
    ref = DC.require_reference()
    import phynetpy._mcmc_seq as mcmc
    import phynetpy._seq_likelihood as rsl
    ref_calc, ref_eng, ref_worker = rsl.FelsensteinCalculator, mcmc._SeqLikelihoodEngine, mcmc._chain_worker
    dropin.install(ref)
    try:
        assert dropin.installed()
        assert mcmc.FelsensteinCalculator is SL.FelsensteinCalculator and rsl.FelsensteinCalculator is SL.FelsensteinCalculator
        assert mcmc._SeqLikelihoodEngine is dropin.B200SeqLikelihoodEngine
        assert mcmc._chain_worker is dropin._chain_worker
        assert mcmc.JC69 is rsl.JC69          # the model classes stay the reference's
    finally:
        dropin.uninstall()
    assert not dropin.installed()
    assert rsl.FelsensteinCalculator is ref_calc and mcmc.FelsensteinCalculator is ref_calc
    assert mcmc._SeqLikelihoodEngine is ref_eng and mcmc._chain_worker is ref_worker


def test_chain_device_policy(monkeypatch):
    monkeypatch.setenv("PHYNET_B200_DEVICES", "0,2,5")
    assert [dropin.chain_device(c) for c in range(7)] == [0, 2, 5, 0, 2, 5, 0]


def test_reference_score_through_the_drop_in(world):
    loci, mapping = DC.make_loci(seed=0, n_loci=4)
    inf = world.mcmc.MCMC_SEQ(loci, mapping, priors=world.mcmc.MCMCSeqPriors(max_reticulations=1))
    got = inf.score()
    dropin.uninstall()
    want = world.mcmc.MCMC_SEQ(loci, mapping, priors=world.mcmc.MCMCSeqPriors(max_reticulations=1)).score()
    dropin.install(world.ref)
    assert math.isfinite(want) and DC.close(got, want)
    assert any(c[0] == "eval" and c[1] == 4 for c in world.dev.calls)   # all four loci in ONE device call


def test_reference_search_runs_on_the_drop_in(world):
    """MCMC_SEQ.search (src/_mcmc_seq.py:3369) unmodified: every proposal is one pending evaluation, accept commits
    on the next evaluation, undo() rolls back; the fake device raises on any violation of the ABI's state rules and
    the chain would then degenerate (propose swallows exceptions into rejects, :2878-2881)."""
    loci, mapping = DC.make_loci(seed=1, n_loci=3)
    inf = world.mcmc.MCMC_SEQ(loci, mapping, priors=world.mcmc.MCMCSeqPriors(max_reticulations=1))
    res = inf.search(num_iter=300, burn_in=50, sample_freq=25, seed=7)
    assert math.isfinite(res.map_log_posterior) and 0.0 < res.acceptance_rate < 1.0 and len(res.samples) == 10
    kinds = [c[0] for c in world.dev.calls]
    assert kinds.count("commit") > 5 and kinds.count("rollback") > 20
    assert len(world.dev.pending) <= 1


@pytest.mark.parametrize("coupled", [False, True])
def test_lockstep_chain_same_values_same_decisions(world, coupled):
    loci, mapping = DC.make_loci(seed=2, n_loci=3)
    n_iter = 250
    log = DC.run_lockstep_chain(world.mcmc, loci, mapping, n_iter, seed=5, max_reticulations=2, coupled=coupled)
    DC.check_lockstep_log(log, n_iter)
    if coupled:
        assert len(log["candidates"]) > 100     # the coupled moves' direct Felsenstein calls went through the proxy


def test_spawned_chain_workers_install_the_drop_in_and_refuse_to_run_without_a_device(monkeypatch):
    """run_parallel_chains (src/_mcmc_seq.py:4076-4274) pickles the sampler into spawned workers.  Each worker must
    come up with the drop-in installed (not silently on the reference's CPU path) -- shown here, without a GPU, by
    both chains failing in pnb_ctx_create with the 'no CPU fallback' message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by tests/test_gpu_dropin.py")
    ref = DC.require_reference()
    import phynetpy._mcmc_seq as mcmc
    monkeypatch.setenv("PHYNETPY_B200_PRELOAD", "baseline.ref_loader:load_reference")
    monkeypatch.setenv("PYTHONPATH", DC.ROOT)
    loci, mapping = DC.make_loci(seed=8, n_loci=2)
    inf = mcmc.MCMC_SEQ(loci, mapping, priors=mcmc.MCMCSeqPriors(max_reticulations=1))
    dropin.install(ref)
    try:
        res = mcmc.run_parallel_chains(inf, n_chains=2, num_iter=20, burn_in=5, sample_freq=5, seed=11,
                                       check_every=10, progress=False)
    finally:
        dropin.uninstall()
    assert sorted(res.errors) == [0, 1]
    assert all("no CPU fallback" in tb for tb in res.errors.values())
