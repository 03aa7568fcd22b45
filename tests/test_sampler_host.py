"""The flat single-chain driver (phynetpy_b200/sampler.py) on the CPU: the real SeqLikelihoodEngine over the fake,
oracle-backed device of test_engine_host.py (which enforces the ABI's pending / commit / rollback rules), so the
whole loop -- operators, prior, accept / reject bookkeeping, caches -- runs without a GPU.  Mirrors the reference's
driver smoke test (tests/test_mcmc_seq.py::TestDriver: finite MAP, acceptance in (0, 1], the strongly supported clade
recovered)."""
import math

import numpy as np
import pytest

from oracle import felsenstein_oracle as O, msnc_oracle as MO
from phynetpy_b200 import _engine as E, netmoves as NM
from phynetpy_b200._seq_likelihood import JC69
from phynetpy_b200.priors import SeqPriors
from phynetpy_b200.sampler import FlatMCMCSeq, OPERATORS, species_tree_from_concatenation
from test_engine_host import FakeDevice, _stub_calc


def _make_data(seed=0, n=150):
    rng = np.random.default_rng(seed)

    def rand(k):
        return "".join(rng.choice(list("ACGT"), size=k))

    def mutate(s, k):
        s = list(s)
        for _ in range(k):
            s[rng.integers(len(s))] = rng.choice(list("ACGT"))
        return "".join(s)

    def locus():
        core = rand(n)
        a = mutate(core, 4)
        return {"A": a, "B": mutate(a, 4), "C": mutate(core, 30)}      # B close to A, C diverged
    return [locus(), locus()], {"A": ["A"], "B": ["B"], "C": ["C"]}


@pytest.fixture
def chain(monkeypatch):
    loci, mapping = _make_data()
    om = O.jc69()
    monkeypatch.setattr(E, "_model_handle", lambda model, ctx: None)

    def fake_msnc(net_, off, parent, blen, ids, theta, ctx=None, branch_thetas=None):
        d = net_.to_dict()
        out = np.empty(len(off) - 1)
        for j in range(len(off) - 1):
            a, b = int(off[j]), int(off[j + 1])
            labels = [net_.species[s] if s >= 0 else None for s in ids[a:b].tolist()]
            out[j] = MO.msnc_log_density(d, parent[a:b].tolist(), blen[a:b].tolist(), labels, {sp: sp for sp in net_.species}, theta)
        return out
    monkeypatch.setattr(E, "msnc_eval_packed", fake_msnc)

    def fake_batch(trees, net_, species_of, theta, *, branch_thetas=None, context=None):
        d = net_.to_dict()
        return np.array([MO.msnc_log_density(d, t.parent.tolist(), t.blen.tolist(), t.label, species_of, theta) for t in trees])
    monkeypatch.setattr(NM, "msnc_log_density_batch", fake_batch)
    dev = FakeDevice(loci, om)
    model = JC69()
    eng = E.SeqLikelihoodEngine([_stub_calc(i, d) for i, d in enumerate(loci)], {"A": "A", "B": "B", "C": "C"}, model, context=dev)
    return FlatMCMCSeq(loci, mapping, priors=SeqPriors(max_reticulations=1), model=model, engine=eng), dev


def _clusters(net):
    kids = [[] for _ in range(net.n_nodes)]
    for s, d in zip(net.edge_src.tolist(), net.edge_dst.tolist()):
        kids[s].append(d)
    memo = {}

    def below(v):
        if v not in memo:
            memo[v] = frozenset([net.leaf_label[v]]) if not kids[v] else frozenset().union(*[below(c) for c in kids[v]])
        return memo[v]
    return {below(v) for v in range(net.n_nodes) if kids[v]}


def test_runs_and_recovers_the_supported_clade(chain):
    inf, dev = chain
    start = inf.score()
    assert math.isfinite(start)
    res = inf.search(num_iter=900, burn_in=200, sample_freq=50, seed=7)
    assert math.isfinite(res.map_log_posterior) and res.map_log_posterior >= start
    assert 0.0 < res.acceptance_rate <= 1.0
    assert len(res.samples) == 14 and all(math.isfinite(s.log_posterior) for s in res.samples)
    assert frozenset({"A", "B"}) in _clusters(res.map_network)
    assert sum(c[0] for c in res.operator_counts.values()) == 900
    assert {n for n, _ in OPERATORS} == set(res.operator_counts)
    assert res.operator_counts["gene_node_height"][1] > 0 and res.operator_counts["change_theta"][1] > 0
    # the chain's cached state is the true state: a from-scratch evaluation of the final state agrees
    final = inf.score()
    for i in range(len(inf.trees)):
        inf.engine.invalidate_locus(i)
    inf.engine.invalidate_network()
    assert inf.score() == pytest.approx(final, rel=1e-12)
    assert all(s.num_reticulations <= 1 for s in res.samples)


def test_species_tree_start_and_bad_inputs(chain):
    inf, dev = chain
    st = species_tree_from_concatenation(inf.loci, inf.mapping)
    assert sorted(st.species) == ["A", "B", "C"] and st.n_reticulations() == 0
    assert frozenset({"A", "B"}) in _clusters(st)
    with pytest.raises(ValueError, match="finite and positive"):
        FlatMCMCSeq(inf.loci, inf.mapping, theta=-1.0, engine=inf.engine)


def test_mc3_ladder_in_lockstep_shares_device_calls(chain, monkeypatch):
    """Two tempered chains: every iteration scores the dirty loci of BOTH chains in one pending evaluation
    (log_likelihood_many), swaps migrate states together with their engines, the cold chain is the one reported."""
    inf, dev = chain
    second = E.SeqLikelihoodEngine([_stub_calc(100 + i, d) for i, d in enumerate(inf.loci)], inf.species_of, inf.model, context=dev)
    dev.loci = dev.loci + [None] * (100 - len(dev.loci)) + [O.OracleLocus(d) for d in inf.loci]
    start = inf.score()
    n_iter = 400
    res = inf.search_mc3([1.0, 2.5], engines=[second], num_iter=n_iter, burn_in=100, sample_freq=50, swap_interval=2, seed=11)
    assert math.isfinite(res.map_log_posterior) and res.map_log_posterior >= start
    assert 0.0 < res.acceptance_rate <= 1.0 and len(res.samples) == 6
    swaps, swaps_ok = res.operator_counts.pop("__swaps__")
    assert swaps == n_iter // 2 and 0 < swaps_ok <= swaps
    assert sum(c[0] for c in res.operator_counts.values()) == n_iter          # cold-chain proposals
    assert frozenset({"A", "B"}) in _clusters(res.map_network)
    # pending evaluations: at most ONE per iteration although two chains stepped (plus the initial one); the other
    # evaluations are the score-only (NOSTORE) batches of the coupled moves
    pend = [c for c in dev.calls if c[0] == "eval" and c[2] & 0x2]
    assert len(pend) <= n_iter + 1
    assert any(c[1] >= 2 for c in pend)                                        # both chains' loci in one call
    # the cold chain's cached state is its true state
    final = inf.score()
    for i in range(len(inf.trees)):
        inf.engine.invalidate_locus(i)
    inf.engine.invalidate_network()
    assert inf.score() == pytest.approx(final, rel=1e-12)
    assert inf.search_mc3([1.0], num_iter=5, burn_in=0, sample_freq=1, seed=1).num_iterations == 5      # one temperature = plain chain
