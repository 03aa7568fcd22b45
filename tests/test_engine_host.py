"""Host logic of SeqLikelihoodEngine on the CPU: the two device calls (``DeviceContext.eval`` / ``commit`` /
``rollback`` and ``msnc_eval_packed``) are replaced by a fake device that (a) answers with the oracles and
(b) enforces the C ABI's state rules -- a locus with a pending evaluation may not be evaluated again before it is
committed or rolled back, nothing may be committed twice.  What is tested is the bookkeeping around the calls:
dirty tracking, commit on the next evaluation / rollback on restore_caches, the packed tree store, the
multi-engine call, totals.  (The same scenarios run against the real kernels in the ``-m gpu`` tests.)"""
import math
import types

import numpy as np
import pytest

from oracle import felsenstein_oracle as O, msnc_oracle as MO
from phynetpy_b200 import _engine as E, _lib, moves as MV, simulate as S
from phynetpy_b200._seq_likelihood import GTR, FelsensteinCalculator
from phynetpy_b200.candidates import node_heights


class FakeDevice:
    """Stands in for DeviceContext: loci are integers, values come from the Felsenstein oracle."""

    def __init__(self, loci_data, model):
        self.loci = [O.OracleLocus(d) for d in loci_data]
        self.model = model
        self.pending = set()
        self.committed_tree = {}
        self.pending_tree = {}
        self.calls = []
        self._h = object()

    def eval(self, model_handle, handles, node_off, parent, blen, tip_row, site_rate=1.0, flags=0, P_edge=None):
        out = np.empty(len(handles))
        nostore = bool(flags & _lib.EVAL_NOSTORE)
        seen = set()
        for j, h in enumerate(handles.tolist()):
            if not nostore:
                assert h not in self.pending, "locus %d evaluated while an evaluation is pending (PNB_ERR_STATE)" % h
                assert h not in seen, "a locus may appear twice in a batch only with NOSTORE"
                seen.add(h)
            a, b = int(node_off[j]), int(node_off[j + 1])
            locus = self.loci[h]
            labels = [None] * (b - a)
            for i, r in enumerate(tip_row[a:b].tolist()):
                if r >= 0 and not (parent[a:b] == i).any():
                    labels[i] = locus.taxa[r]
            tree = O.OracleTree(parent[a:b].tolist(), [None if x != x else x for x in blen[a:b].tolist()], labels)
            out[j] = O.log_likelihood(locus, tree, self.model)
            if not nostore:
                key = (parent[a:b].tobytes(), blen[a:b].tobytes())
                if flags & _lib.EVAL_PENDING:
                    self.pending.add(h)
                    self.pending_tree[h] = key
                else:
                    self.committed_tree[h] = key
        self.calls.append(("eval", len(handles), int(flags)))
        return out

    def commit(self, handles):
        for h in handles.tolist():
            assert h in self.pending, "commit of a locus without a pending evaluation"
            self.pending.discard(h)
            self.committed_tree[h] = self.pending_tree.pop(h)
        self.calls.append(("commit", len(handles)))

    def rollback(self, handles):
        for h in handles.tolist():
            assert h in self.pending, "rollback of a locus without a pending evaluation"
            self.pending.discard(h)
            self.pending_tree.pop(h)
        self.calls.append(("rollback", len(handles)))

    def last_stats(self):
        return {"launches": 1, "updates": 0, "nodes_recomputed": 0}


def _stub_calc(i, data):
    c = FelsensteinCalculator.__new__(FelsensteinCalculator)
    c.taxa = list(data)
    c._row_of = {t: k for k, t in enumerate(c.taxa)}
    c._locus_handle = lambda i=i: types.SimpleNamespace(value=i)
    return c


@pytest.fixture
def world(monkeypatch):
    rng = np.random.default_rng(17)
    n_species, M, L = 6, 8, 60
    net = S.random_species_network(n_species, 1, rng, height=0.03)
    trees, species_of = S.simulate_gene_trees_msnc(net, 1, 0.02, rng, M)
    aln = S.simulate_alignments(np.stack([t.parent for t in trees]), np.stack([t.blen for t in trees]), L,
                                S.DEFAULT_PI, S.DEFAULT_RATES, rng)
    data = [S.to_alignment_dict(aln[i]) for i in range(M)]
    om = O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES)
    monkeypatch.setattr(E, "_model_handle", lambda model, ctx: None)
    netd = net.to_dict()
    msnc_calls = []

    def fake_msnc(net_, off, parent, blen, ids, theta, ctx=None, branch_thetas=None):
        out = np.empty(len(off) - 1)
        for j in range(len(off) - 1):
            a, b = int(off[j]), int(off[j + 1])
            labels = [net_.species[s] if s >= 0 else None for s in ids[a:b].tolist()]     # one allele per species, same names
            out[j] = MO.msnc_log_density(netd, parent[a:b].tolist(), blen[a:b].tolist(), labels,
                                         {sp: sp for sp in net_.species}, theta)
        msnc_calls.append(len(off) - 1)
        return out
    monkeypatch.setattr(E, "msnc_eval_packed", fake_msnc)

    def make_engine():
        dev = FakeDevice(data, om)
        eng = E.SeqLikelihoodEngine([_stub_calc(i, d) for i, d in enumerate(data)], species_of, GTR(S.DEFAULT_PI, S.DEFAULT_RATES),
                                    context=dev)
        return eng, dev
    # allele t_k maps to species s_k: the fake MSNC labels leaves by species, so rename through species_of
    return types.SimpleNamespace(net=net, netd=netd, trees=trees, species_of=species_of, data=data, om=om, M=M,
                                 make_engine=make_engine, msnc_calls=msnc_calls, rng=rng)


def _expected_total(w, trees, theta):
    tot = 0.0
    f = [O.log_likelihood(O.OracleLocus(d), O.OracleTree(t.parent.tolist(), [None if x != x else x for x in t.blen.tolist()], t.label), w.om)
         for d, t in zip(w.data, trees)]
    m = [MO.msnc_log_density(w.netd, t.parent.tolist(), t.blen.tolist(), t.label, w.species_of, theta) for t in trees]
    return sum(f) + sum(m), f, m


def test_chain_bookkeeping_against_the_state_rules(world):
    w = world
    eng, dev = w.make_engine()
    theta = 0.02
    cur = list(w.trees)
    heights = [node_heights(t) for t in cur]
    total = eng.log_likelihood(cur, w.net, theta)
    want, f, m = _expected_total(w, cur, theta)
    assert total == pytest.approx(want, rel=1e-13)
    assert dev.calls[0][:2] == ("eval", w.M) and dev.calls[0][2] & _lib.EVAL_PENDING and w.msnc_calls == [w.M]
    rng = np.random.default_rng(5)
    n_acc = n_rej = 0
    for step in range(60):
        snap = eng.clone_caches()
        if step % 7 == 6:                                   # theta move: all MSNC entries, no pruning
            new_theta = theta * 1.1
            eng.invalidate_network()
            n_eval = sum(c[0] == "eval" for c in dev.calls)
            new_total = eng.log_likelihood(cur, w.net, new_theta)
            assert sum(c[0] == "eval" for c in dev.calls) == n_eval          # nothing dirty on the Felsenstein side
            assert w.msnc_calls[-1] == w.M
            prop, prop_h = cur, None
        else:
            new_theta = theta
            i = int(rng.integers(w.M))
            mv = MV.slide_height_move(cur[i], heights[i], rng) if step % 2 else None
            if mv is None:
                out = MV.nni_move(cur[i], heights[i], rng)
                if out is None:
                    continue
                mv = (out[0], out[1], 0.0)
            prop = list(cur)
            prop[i] = mv[0]
            prop_h = (i, mv[1])
            eng.invalidate_locus(i)
            new_total = eng.log_likelihood(prop, w.net, theta)
            assert dev.calls[-1][:2] == ("eval", 1) or dev.calls[-2][:2] == ("eval", 1)
            assert w.msnc_calls[-1] == 1                                      # one density for one dirty locus
        assert dev.pending == ({prop_h[0]} if prop_h else set())
        if rng.random() < 0.5:
            cur, total, theta = prop, new_total, new_theta
            if prop_h:
                heights[prop_h[0]] = prop_h[1]
            n_acc += 1
        else:
            eng.restore_caches(snap)
            assert not dev.pending                                            # the rejected evaluation was rolled back
            n_rej += 1
    assert n_acc > 10 and n_rej > 10
    final = eng.log_likelihood(cur, w.net, theta)                             # commits whatever was accepted last
    assert not dev.pending
    want, f, m = _expected_total(w, cur, theta)
    assert final == pytest.approx(want, rel=1e-13)
    for a, b in zip(eng._felsen, f):
        assert a == pytest.approx(b, rel=1e-13)
    for a, b in zip(eng._msnc, m):
        assert (a == b) or a == pytest.approx(b, rel=1e-13)
    # the device's committed trees are the accepted ones
    for i, t in enumerate(cur):
        assert dev.committed_tree[i] == (t.parent.tobytes(), t.blen.tobytes())


def test_entries_invalidated_behind_the_engines_back_are_recounted(world):
    w = world
    eng, dev = w.make_engine()
    total = eng.log_likelihood(w.trees, w.net, 0.02)
    eng._felsen[3] = None                       # a caller sharing the list
    eng._msnc[3] = None
    eng._msnc[5] = None
    assert eng.log_likelihood(w.trees, w.net, 0.02) == pytest.approx(total, rel=1e-14)
    assert dev.calls[-1][:2] in (("eval", 1), ("commit", 1)) and w.msnc_calls[-1] == 2


def test_several_engines_share_one_call_and_stay_pending(world):
    w = world
    (e1, d1) = w.make_engine()
    e2 = E.SeqLikelihoodEngine([_stub_calc(i + 100, d) for i, d in enumerate(w.data)], w.species_of, e1._model, context=d1)
    d1.loci = d1.loci + [None] * (100 - len(d1.loci)) + [O.OracleLocus(d) for d in w.data]     # handles 100.. = second chain
    t2 = [t.with_lengths(t.blen * 1.05) for t in w.trees]
    many = E.SeqLikelihoodEngine.log_likelihood_many([e1, e2], [w.trees, t2], [w.net, w.net], [0.02, 0.02])
    evals = [c for c in d1.calls if c[0] == "eval"]
    assert len(evals) == 1 and evals[0][1] == 2 * w.M                         # ONE device call for both chains
    assert d1.pending == set(range(w.M)) | set(range(100, 100 + w.M))         # ... and everything still pending
    assert many[0] == pytest.approx(_expected_total(w, w.trees, 0.02)[0], rel=1e-13)
    assert many[1] == pytest.approx(_expected_total(w, t2, 0.02)[0], rel=1e-13)
    # chain 1 moves one locus and rejects, chain 2 moves one locus and accepts
    s1, s2 = e1.clone_caches(), e2.clone_caches()
    p1, p2 = list(w.trees), list(t2)
    p1[2] = w.trees[2].with_lengths(w.trees[2].blen * 1.2)
    p2[4] = t2[4].with_lengths(t2[4].blen * 0.9)
    e1.invalidate_locus(2)
    e2.invalidate_locus(4)
    many2 = E.SeqLikelihoodEngine.log_likelihood_many([e1, e2], [p1, p2], [w.net, w.net], [0.02, 0.02])
    assert d1.pending == {2, 104}
    assert [c for c in d1.calls if c[0] == "eval"][-1][1] == 2
    e1.restore_caches(s1)
    assert d1.pending == {104}
    many3 = E.SeqLikelihoodEngine.log_likelihood_many([e1, e2], [w.trees, p2], [w.net, w.net], [0.02, 0.02])
    assert not d1.pending and many3[0] == many[0] and many3[1] == many2[1]
    with pytest.raises(ValueError, match="one gene-tree list per engine"):
        E.SeqLikelihoodEngine.log_likelihood_many([e1, e2], [w.trees])


def test_network_only_msnc_score_is_the_placement_scan_sum(world):
    """``network_only_msnc_score`` (the likelihood part of the reference's ``_network_only_log_score``,
    ``src/_mcmc_seq.py:1163-1214``): sum over loci of the log-sum-exp over the height-scale grid of the density of the
    scaled gene tree -- one packed device call per candidate network; the pack is reused while the gene-tree OBJECTS,
    the species and the grid stay the same, and the candidate never touches the engine's own network / cache state."""
    from phynetpy_b200 import netmoves as NM
    w = world
    eng, dev = w.make_engine()
    theta = 0.02
    grid = (0.25, 0.5, 1.0, 2.0, 4.0)
    total0 = eng.log_likelihood(list(w.trees), w.net, theta)
    felsen0, msnc0 = list(eng._felsen), list(eng._msnc)
    n_calls = len(w.msnc_calls)

    def want(trees, net):
        rows = []
        d = net.to_dict()
        for t in trees:
            for s in grid:
                rows.append(MO.msnc_log_density(d, t.parent.tolist(), (t.blen * s).tolist(), t.label, w.species_of, theta))
        return NM.network_only_msnc_score(np.asarray(rows), len(grid))

    got = eng.network_only_msnc_score(w.trees, w.net, theta, grid)
    assert w.msnc_calls[n_calls:] == [w.M * len(grid)]                   # ONE call over loci x scales
    exp = want(w.trees, w.net)
    assert (got == exp == -math.inf) or got == pytest.approx(exp, rel=1e-12)
    pack = eng._scan_pack
    # a candidate network (heights moved): same gene-tree objects -> the pack is reused
    cand = None
    for _ in range(20):
        out = MV.net_slide_height_move(w.net, w.rng)
        if out is not None:
            cand = out[0]
            break
    assert cand is not None
    got2 = eng.network_only_msnc_score(w.trees, cand, theta, grid)
    assert eng._scan_pack is pack
    exp2 = want(w.trees, cand)
    assert (got2 == exp2 == -math.inf) or got2 == pytest.approx(exp2, rel=1e-12)
    # one gene-tree object replaced -> repacked; a different grid -> repacked
    trees2 = list(w.trees)
    trees2[1] = trees2[1].with_lengths(trees2[1].blen * 1.5)
    got3 = eng.network_only_msnc_score(trees2, w.net, theta, grid)
    assert eng._scan_pack is not pack
    exp3 = want(trees2, w.net)
    assert (got3 == exp3 == -math.inf) or got3 == pytest.approx(exp3, rel=1e-12)
    pack3 = eng._scan_pack
    eng.network_only_msnc_score(trees2, w.net, theta, (0.5, 1.0, 2.0))
    assert eng._scan_pack is not pack3 and w.msnc_calls[-1] == w.M * 3
    # the engine's own state was never touched
    assert eng._felsen == felsen0 and eng._msnc == msnc0
    assert eng.log_likelihood(list(w.trees), w.net, theta) == total0


def test_summed_list_total_is_a_function_of_the_content():
    """The cache lists keep block sums: the total after any sequence of writes equals the total of a fresh list with the
    same content BIT FOR BIT (cached == from-scratch), a None anywhere raises TypeError as sum() would, copies carry
    their sums, every other list mutation stays correct, and the list pickles as a list."""
    import pickle
    rng = np.random.default_rng(3)
    n = 1000
    vals = (-rng.random(n) * 1e4).tolist()
    a = E._SummedList(vals)
    t0 = a.total()
    assert t0 == E._SummedList(list(vals)).total()
    assert abs(t0 - sum(vals)) <= 1e-9 * abs(t0)
    for _ in range(200):
        i = int(rng.integers(-n, n))
        a[i] = float(-rng.random() * 1e4)
        if rng.random() < 0.3:
            assert a.total() == E._SummedList(list(a)).total()
    b = a.copy()
    assert type(b) is E._SummedList and b == a and b.total() == a.total()
    b[5] = 1.0
    assert a[5] != 1.0 and a.total() == E._SummedList(list(a)).total()
    a[10:20] = [0.5] * 10                       # slice assignment (the engine's contiguous-run path)
    assert a.total() == E._SummedList(list(a)).total()
    a[777] = None                               # anybody, through any reference
    with pytest.raises(TypeError):
        a.total()
    a[777] = -2.0
    a.append(-1.0)
    a.pop(0)
    assert a.total() == E._SummedList(list(a)).total()
    c = pickle.loads(pickle.dumps(a))
    assert type(c) is E._SummedList and list(c) == list(a) and c.total() == a.total()
    small = E._SummedList([1.0, 2.0, None])
    with pytest.raises(TypeError):
        small.total()
    assert E._SummedList([-math.inf, 1.0] * 40).total() == -math.inf
    assert E._list_total([1.0, 2.0]) == 3.0 and type(E._copy_cache([1.0])) is E._SummedList
