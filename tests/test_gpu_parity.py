"""GPU parity tests: the CUDA path, through the C ABI, against (a) the golden vectors
frozen from the reference and (b) the CPU oracle on seeded inputs.

Bars: integers (pattern count, order, weights, tip codes) bit-exact; log-likelihood
within 1e-10 relative (north_star) -- in practice ~1e-14; P(t) within 1e-13 abs.
"""
import math

import numpy as np
import pytest

from oracle import felsenstein_oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu

CASES = load_golden("felsen_cases.json")["cases"]
PM = load_golden("pmatrix_cases.json")["cases"]
REL = 1e-10


def _pkg():
    from phynetpy_b200 import infer as P
    return P


def _model(P, spec):
    if spec["kind"] == "JC69":
        return P.JC69()
    if spec["kind"] == "HKY85":
        return P.HKY85(spec["kappa"], spec["pi"])
    return P.GTR(spec["pi"], spec["rates"])


def _omodel(spec):
    if spec["kind"] == "JC69":
        return O.jc69()
    if spec["kind"] == "HKY85":
        return O.hky85(spec["kappa"], spec["pi"])
    return O.gtr(spec["pi"], spec["rates"])


def _tree(P, t):
    return P.GeneTree(t["parent"], t["blen"], t["label"])


def _otree(ft):
    return O.OracleTree(ft.parent.tolist(), [None if math.isnan(b) else float(b) for b in ft.blen], ft.label)


def rel_err(a, b):
    return abs(a - b) / max(1.0, abs(b))


# ----------------------------------------------------------------- golden vectors
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_logL_and_patterns(case):
    P = _pkg()
    fc = P.FelsensteinCalculator(dict(case["alignment"]))
    assert fc.n_patterns == case["n_patterns"]
    assert fc._int_weights.tolist() == case["weights"]
    got = fc.log_likelihood(_tree(P, case["tree"]), _model(P, case["model"]), site_rate=case["site_rate"])
    assert rel_err(got, case["logL"]) <= REL, (got, case["logL"])
    # a second call is served from the device cache and must return the same bits
    again = fc.log_likelihood(_tree(P, case["tree"]), _model(P, case["model"]), site_rate=case["site_rate"])
    assert again == got
    full = fc.log_likelihood(_tree(P, case["tree"]), _model(P, case["model"]), site_rate=case["site_rate"], full=True)
    assert full == got


@pytest.mark.parametrize("case", CASES[:30], ids=[c["name"] for c in CASES[:30]])
def test_device_compress_bit_exact(case):
    P = _pkg()
    from phynetpy_b200 import _seq_likelihood as SL
    seqs = [s for _, s in case["alignment"]]
    mat, lut = SL._alignment_matrix(seqs)
    cd, wd, fd, pd_ = SL.compress_alignment(mat, lut, use_device=True)
    ch, wh, fh, ph = SL.compress_alignment(mat, lut, use_device=False)
    assert wd.tolist() == case["weights"]
    assert np.array_equal(cd, ch) and np.array_equal(wd, wh) and np.array_equal(fd, fh) and np.array_equal(pd_, ph)


@pytest.mark.parametrize("case", PM, ids=["%s_%d" % (c.get("cls", c.get("model", {}).get("kind")), i) for i, c in enumerate(PM)])
def test_pmatrix_golden(case):
    P = _pkg()
    if case["family"] == "_seq_likelihood":
        m = _model(P, case["model"])
        got = m.p_matrix_batch(case["times"])
        single = np.stack([m.p_matrix(t) for t in case["times"][:3]])
        assert np.array_equal(single, got[:3])
    else:
        from phynetpy_b200 import GTR as G
        m = getattr(G, case["cls"])(*case["args"])
        np.testing.assert_allclose(np.asarray(m.getQ()).ravel(), case["Q"], atol=1e-15)
        got = m.expt_batch(case["times"])
        np.testing.assert_allclose(m.expt(case["times"][5]), got[5], atol=0)
    exp = np.asarray(case["P"]).reshape(-1, 4, 4)
    # reference tests pin expt/p_matrix to expm at 1e-11 (tests/test_gtr.py:221-232, :498-533)
    np.testing.assert_allclose(got, exp, rtol=0, atol=1e-13)


def test_pmatrix_properties():
    """Properties the reference tests (tests/test_gtr.py:234-285): stochastic rows, P(0)=I,
    P(500)->pi, Chapman-Kolmogorov, negative-t clamp, detailed balance."""
    P = _pkg()
    m = P.GTR([0.1, 0.2, 0.3, 0.4], [1.0, 2.5, 0.7, 1.3, 3.1, 0.9])
    ts = np.array([0.0, 1e-4, 0.01, 0.1, 0.3, 0.4, 0.7, 1.0, 50.0, 500.0, -1.0])
    Pm = m.p_matrix_batch(ts)
    np.testing.assert_allclose(Pm.sum(axis=2), 1.0, atol=1e-13)
    np.testing.assert_allclose(Pm[0], np.eye(4), atol=1e-14)
    np.testing.assert_allclose(Pm[-1], np.eye(4), atol=1e-14)
    np.testing.assert_allclose(Pm[9], np.tile(m.pi, (4, 1)), atol=1e-8)
    np.testing.assert_allclose(Pm[4] @ Pm[5], Pm[6], atol=1e-13)
    flux = m.pi[:, None] * Pm[3]
    np.testing.assert_allclose(flux, flux.T, atol=1e-14)
    assert (Pm >= -1e-15).all()


# ------------------------------------------------------------- oracle on seeded inputs
@pytest.mark.parametrize("n,L,seed", [(5, 1000, 1), (16, 500, 4), (24, 1000, 3), (32, 2000, 5), (50, 3000, 2)])
def test_baseline_shapes_vs_oracle(n, L, seed):
    """BASELINE.json shapes at sizes the oracle finishes in seconds (cfg1 exactly; cfg2-5 one locus)."""
    P = _pkg()
    from phynetpy_b200 import simulate as S
    rng = np.random.default_rng(seed)
    par, bl = S.random_trees(1, n, rng)
    aln = S.simulate_alignments(par, bl, L, S.DEFAULT_PI, S.DEFAULT_RATES, rng)[0]
    ft = S.flat_trees(par, bl)[0]
    fc = P.FelsensteinCalculator(S.to_alignment_dict(aln))
    locus = O.OracleLocus(S.to_alignment_dict(aln))
    assert fc.n_patterns == locus.n_patterns and fc._int_weights.tolist() == locus.int_weights.tolist()
    assert np.array_equal(fc._codes, locus.codes())
    got = fc.log_likelihood(ft, P.GTR(S.DEFAULT_PI, S.DEFAULT_RATES))
    sink = {}
    exp = O.log_likelihood(locus, _otree(ft), O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES), sink=sink)
    assert rel_err(got, exp) <= REL
    # every internal node's partial, as stored in HBM, against the oracle's
    # tip-only nodes ("cherries") are recomputed from the 1-byte codes instead of being kept in HBM
    nc = ft.n_children()
    tips = (nc == 0).astype(int)
    for v in range(ft.n_nodes):          # random_trees numbers children before parents
        if ft.parent[v] >= 0:
            tips[ft.parent[v]] += tips[v]
    checked = 0
    for node in np.nonzero(nc > 0)[0]:
        kids = np.nonzero(ft.parent == node)[0]
        try:
            got_partial = fc.node_partial(int(node))
        except P.PhyNetB200Error as e:
            # only roots of small subtrees (tip-only, or <= PNB_VIRT_TIPS = 4 tips) may be absent (one that had to be
            # held for a sibling IS materialised)
            assert "small subtree" in str(e) and ((nc[kids] == 0).all() or tips[node] <= 4)
            continue
        np.testing.assert_allclose(got_partial, sink[int(node)], rtol=1e-12, atol=0)
        checked += 1
    assert checked >= (n - 1) // 4 or n <= 8


def test_dirty_path_equals_full_recompute():
    """K-c: after a one-branch change only the path to the root is recomputed, and the
    result is bit-identical to a full recomputation and within 1e-10 of the oracle."""
    P = _pkg()
    from phynetpy_b200 import simulate as S
    rng = np.random.default_rng(44)
    n, L = 16, 500
    par, bl = S.random_trees(1, n, rng)
    aln = S.simulate_alignments(par, bl, L, S.DEFAULT_PI, S.DEFAULT_RATES, rng)[0]
    ft = S.flat_trees(par, bl)[0]
    model = P.GTR(S.DEFAULT_PI, S.DEFAULT_RATES)
    fc = P.FelsensteinCalculator(S.to_alignment_dict(aln))
    ctx = fc._context()
    base = fc.log_likelihood(ft, model)
    assert ctx.last_stats()["nodes_recomputed"] == n - 1
    locus = O.OracleLocus(S.to_alignment_dict(aln))
    om = O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES)
    for step in range(12):
        node = int(rng.integers(0, 2 * n - 2))
        bl2 = ft.blen.copy()
        bl2[node] *= float(rng.uniform(0.5, 2.0))
        ft2 = ft.with_lengths(bl2)
        got = fc.log_likelihood(ft2, model)
        st = ctx.last_stats()
        depth = 0
        p = ft.parent[node]
        while p >= 0:
            depth += 1
            p = ft.parent[p]
        assert st["nodes_recomputed"] == depth, (st, depth)
        assert st["updates"] == depth * fc.n_patterns
        exp = O.log_likelihood(locus, _otree(ft2), om)
        assert rel_err(got, exp) <= REL
        fc2 = P.FelsensteinCalculator(S.to_alignment_dict(aln))
        assert fc2.log_likelihood(ft2, model) == got  # bit-identical to a cold full evaluation
        fc2.close()
        ft = ft2
    assert base != got


def test_nni_topology_change_dirty_path():
    P = _pkg()
    from phynetpy_b200 import simulate as S
    rng = np.random.default_rng(9)
    n = 12
    par, bl = S.random_trees(1, n, rng)
    aln = S.simulate_alignments(par, bl, 400, S.DEFAULT_PI, S.DEFAULT_RATES, rng)[0]
    ft = S.flat_trees(par, bl)[0]
    model = P.HKY85(3.0, [0.2, 0.3, 0.3, 0.2])
    om = O.hky85(3.0, [0.2, 0.3, 0.3, 0.2])
    fc = P.FelsensteinCalculator(S.to_alignment_dict(aln))
    locus = O.OracleLocus(S.to_alignment_dict(aln))
    fc.log_likelihood(ft, model)
    parent = ft.parent.copy()
    # NNI around an internal edge (v -> u): swap a child of v with v's sibling
    internal = [i for i in range(n, 2 * n - 2)]
    v = internal[0]
    u = int(parent[v])
    child = int(np.nonzero(parent == v)[0][0])
    sib = [int(c) for c in np.nonzero(parent == u)[0] if c != v][0]
    parent[child], parent[sib] = u, v
    ft2 = P.FlatTree(parent, ft.blen, ft.label)
    got = fc.log_likelihood(ft2, model)
    assert fc._context().last_stats()["nodes_recomputed"] < n - 1
    exp = O.log_likelihood(locus, _otree(ft2), om)
    assert rel_err(got, exp) <= REL


def test_engine_golden_and_cache_semantics(engine_case):
    P = _pkg()
    ec = engine_case
    model = _model(P, ec["model"])
    loci = [dict(l["alignment"]) for l in ec["loci"]]
    trees = [_tree(P, l["tree"]) for l in ec["loci"]]
    eng = P.SeqLikelihoodEngine(loci, None, model)
    total = eng.log_likelihood(trees, None, 0.0)
    assert rel_err(total, ec["total"]) <= REL
    for got, exp in zip(eng._felsen, ec["per_locus"]):
        assert rel_err(got, exp) <= REL
    # reject path: snapshot, perturb one locus, evaluate, restore -> old total, device rolled back
    snap = eng.clone_caches()
    i = 3
    ft = trees[i].flat()
    bl = ft.blen.copy()
    leaf = int(np.nonzero(ft.n_children() == 0)[0][0])
    bl[leaf] *= 3.0
    new_trees = list(trees)
    new_trees[i] = ft.with_lengths(bl)
    eng.invalidate_locus(i)
    t2 = eng.log_likelihood(new_trees, None, 0.0)
    assert t2 != total and eng.last_stats["jobs_cached"] == 0
    assert eng.last_stats["nodes_recomputed"] < (ft.n_nodes - 1)
    eng.restore_caches(snap)
    assert eng.log_likelihood(trees, None, 0.0) == total
    # accept path: evaluate again, keep; then the same tree is a pure cache hit
    eng.invalidate_locus(i)
    t3 = eng.log_likelihood(new_trees, None, 0.0)
    assert t3 == t2
    eng.invalidate_locus(i)
    t4 = eng.log_likelihood(new_trees, None, 0.0)
    assert t4 == t2 and eng.last_stats["jobs_cached"] == 1
    # candidates: many trees for one locus, nothing stored
    cands = [new_trees[i], trees[i], new_trees[i]]
    vals = eng.score_candidates(i, cands)
    assert vals[0] == vals[2]
    assert rel_err(vals[1], ec["per_locus"][i]) <= REL
    # candidate sets for several loci in one launch (the coupled moves' double loop)
    sets = [(j, [trees[j], trees[j].flat().with_lengths(trees[j].flat().blen * 2.0)]) for j in (0, 5, 11)]
    res = eng.score_candidate_sets(sets)
    assert [len(r) for r in res] == [2, 2, 2]
    for (j, _), r in zip(sets, res):
        assert rel_err(r[0], ec["per_locus"][j]) <= REL and r[1] != r[0]
    assert eng.last_stats["launches"] in (1, 2) and eng.log_likelihood(new_trees, None, 0.0) == t2


def test_candidates_nostore_deep_tree_uses_stash():
    """NOSTORE evaluation of a balanced tree needs the shared-memory stash; result == stored path."""
    P = _pkg()
    from phynetpy_b200 import simulate as S
    rng = np.random.default_rng(21)
    n = 32
    # perfectly balanced topology: pair up nodes level by level
    parent = np.full(2 * n - 1, -1, dtype=np.int32)
    level = list(range(n))
    nxt = n
    while len(level) > 1:
        new = []
        for a, b in zip(level[::2], level[1::2]):
            parent[a] = parent[b] = nxt
            new.append(nxt)
            nxt += 1
        level = new
    blen = rng.exponential(0.05, size=2 * n - 1)
    blen[-1] = np.nan
    par2, bl2 = parent[None, :], blen[None, :]
    aln = S.simulate_alignments(par2, bl2, 700, S.DEFAULT_PI, S.DEFAULT_RATES, rng)[0]
    ft = P.FlatTree(parent, blen, S.tip_labels(n) + [None] * (n - 1))
    model = P.GTR(S.DEFAULT_PI, S.DEFAULT_RATES)
    eng = P.SeqLikelihoodEngine([S.to_alignment_dict(aln)], None, model)
    stored = eng.log_likelihood([ft], None, 0.0)
    exp = O.log_likelihood(O.OracleLocus(S.to_alignment_dict(aln)), _otree(ft), O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES))
    assert rel_err(stored, exp) <= REL
    scaled = ft.with_lengths(ft.blen * 1.7)   # every node dirty -> full NOSTORE walk with stash
    vals = eng.score_candidates(0, [scaled, ft])
    fc = P.FelsensteinCalculator(S.to_alignment_dict(aln))
    assert vals[0] == fc.log_likelihood(scaled, model)
    assert vals[1] == stored


def test_explicit_model_subclass_overriding_p_matrix():
    P = _pkg()
    base = O.gtr([0.1, 0.2, 0.3, 0.4], [1.0, 2.5, 0.7, 1.3, 3.1, 0.9])

    class HostModel(P.SubstitutionModel):
        def p_matrix(self, t):  # user code on the host, as the reference allows
            return base.p_matrix(t)

    case = CASES[3]
    fc = P.FelsensteinCalculator(dict(case["alignment"]))
    got = fc.log_likelihood(_tree(P, case["tree"]), HostModel([0.1, 0.2, 0.3, 0.4], [1.0, 2.5, 0.7, 1.3, 3.1, 0.9]))
    assert rel_err(got, case["logL"]) <= REL


def test_model_change_invalidates_cached_partials():
    P = _pkg()
    case = CASES[0]
    fc = P.FelsensteinCalculator(dict(case["alignment"]))
    t = _tree(P, case["tree"])
    a = fc.log_likelihood(t, P.JC69())
    b = fc.log_likelihood(t, P.HKY85(2.5, [0.3, 0.2, 0.2, 0.3]))
    assert rel_err(a, CASES[0]["logL"]) <= REL and rel_err(b, CASES[1]["logL"]) <= REL


def test_edge_cases_empty_and_single_node():
    P = _pkg()
    fc = P.FelsensteinCalculator({"A": "", "B": ""})
    assert fc.n_patterns == 0
    assert fc.log_likelihood(P.GeneTree.from_newick("(A:0.1,B:0.1);"), P.JC69()) == 0.0
    # a tree that is a single leaf: likelihood is pi[state] per site (reference :408-413 + :394-397)
    fc1 = P.FelsensteinCalculator({"A": "ACGTTN"})
    m = P.GTR([0.1, 0.2, 0.3, 0.4], [1] * 6)
    single = P.GeneTree([-1], [None], ["A"])
    exp = math.log(0.1) + math.log(0.2) + math.log(0.3) + 2 * math.log(0.4) + math.log(1.0)
    assert fc1.log_likelihood(single, m) == pytest.approx(exp, rel=1e-13)
    with pytest.raises(ValueError):
        fc1.log_likelihood(P.FlatTree([1, 0], [0.1, 0.1], ["A", None]), m)  # cycle / no root


def test_ragged_batch_many_loci_vs_oracle():
    """Ragged loci (different taxa counts, pattern counts around the 128/512 tile edges) in one batch."""
    P = _pkg()
    from phynetpy_b200 import simulate as S
    rng = np.random.default_rng(77)
    model = P.GTR(S.DEFAULT_PI, S.DEFAULT_RATES)
    om = O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES)
    loci, trees, exp = [], [], []
    for n, L in [(4, 127), (6, 128), (9, 129), (12, 511), (5, 512), (7, 513), (20, 1300), (3, 1), (11, 2500)]:
        par, bl = S.random_trees(1, n, rng, height=0.6)
        aln = S.simulate_alignments(par, bl, L, S.DEFAULT_PI, S.DEFAULT_RATES, rng)[0]
        ft = S.flat_trees(par, bl)[0]
        d = S.to_alignment_dict(aln)
        loci.append(d)
        trees.append(ft)
        exp.append(O.log_likelihood(O.OracleLocus(d), _otree(ft), om))
    eng = P.SeqLikelihoodEngine(loci, None, model)
    total = eng.log_likelihood(trees, None, 0.0)
    for got, e in zip(eng._felsen, exp):
        assert rel_err(got, e) <= REL
    assert rel_err(total, sum(exp)) <= REL


def test_large_locus_full_size_properties():
    """BASELINE cfg2 at full size on one GPU (50 taxa, 1e6 unique patterns): size-independent properties --
    (1) total == sum over disjoint pattern blocks evaluated as separate loci (linearity),
    (2) dirty-path == full, (3) oracle agreement on a 20k-pattern prefix block."""
    P = _pkg()
    from phynetpy_b200 import simulate as S
    rng = np.random.default_rng(2)
    n, L = 50, 1_060_000
    par, bl = S.random_trees(1, n, rng)
    aln = np.concatenate([S.simulate_alignments(par, bl, min(132_500, L - c0), S.DEFAULT_PI, S.DEFAULT_RATES, rng)[0]
                          for c0 in range(0, L, 132_500)], axis=1)
    ft = S.flat_trees(par, bl)[0]
    model = P.GTR(S.DEFAULT_PI, S.DEFAULT_RATES)
    taxa = S.tip_labels(n)
    fc = P.FelsensteinCalculator.from_matrix(taxa, aln, device_compress=True)
    assert fc.n_patterns >= 1_000_000 and int(fc._int_weights.sum()) == L
    whole = fc.log_likelihood(ft, model)
    parts = []
    for lo in range(0, fc.n_patterns, 200_000):
        sub = P.FelsensteinCalculator.from_codes(taxa, fc._codes[:, lo:lo + 200_000], fc._int_weights[lo:lo + 200_000])
        parts.append(sub.log_likelihood(ft, model))
        if lo == 0:
            exp0 = O.log_likelihood_codes(fc._codes[:, :20_000], fc._int_weights[:20_000],
                                          ft.tip_rows({t: i for i, t in enumerate(taxa)}), _otree(ft),
                                          O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES))
            sub0 = P.FelsensteinCalculator.from_codes(taxa, fc._codes[:, :20_000], fc._int_weights[:20_000])
            assert rel_err(sub0.log_likelihood(ft, model), exp0) <= REL
        sub.close()
    assert rel_err(sum(parts), whole) <= 1e-12
    bl2 = ft.blen.copy()
    bl2[7] *= 1.5
    ft2 = ft.with_lengths(bl2)
    dirty = fc.log_likelihood(ft2, model)
    assert fc._context().last_stats()["nodes_recomputed"] < n - 1
    assert dirty == fc.log_likelihood(ft2, model, full=True)


def test_big_batch_chunked_parallel_compile_vs_oracle():
    """A batch large enough for the chunked launch path + multi-threaded tree compilation
    (>= 32768 nodes): per-locus results equal single-job evaluations bit for bit and the oracle
    to 1e-10; then a one-branch change in EVERY locus (dirty path in every job) and a rollback."""
    P = _pkg()
    from phynetpy_b200 import simulate as S
    rng = np.random.default_rng(123)
    M, n, L = 640, 30, 90
    par, bl = S.random_trees(M, n, rng, height=0.4)
    aln = S.simulate_alignments(par, bl, L, S.DEFAULT_PI, S.DEFAULT_RATES, rng)
    trees = S.flat_trees(par, bl)
    taxa = S.tip_labels(n)
    calcs = [P.FelsensteinCalculator.from_matrix(taxa, aln[i], device_compress=False) for i in range(M)]
    model = P.GTR(S.DEFAULT_PI, S.DEFAULT_RATES)
    eng = P.SeqLikelihoodEngine(calcs, None, model)
    total = eng.log_likelihood(trees, None, 0.0)
    assert eng.last_stats["launches"] >= 4  # chunked: several (K-a, prune) pairs
    om = O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES)
    for i in list(range(0, M, 97)) + [M - 1]:
        exp = O.log_likelihood(O.OracleLocus(S.to_alignment_dict(aln[i])), _otree(trees[i]), om)
        assert rel_err(eng._felsen[i], exp) <= REL
        solo = P.FelsensteinCalculator.from_matrix(taxa, aln[i], device_compress=False)
        assert solo.log_likelihood(trees[i], model) == eng._felsen[i]
        solo.close()
    assert rel_err(total, sum(eng._felsen)) <= 1e-13
    # dirty path in every locus at once
    snap = eng.clone_caches()
    new_trees = []
    for i, t in enumerate(trees):
        b = t.blen.copy()
        b[i % n] *= 1.5
        new_trees.append(t.with_lengths(b))
        eng.invalidate_locus(i)
    t2 = eng.log_likelihood(new_trees, None, 0.0)
    assert eng.last_stats["nodes_recomputed"] < M * (n - 1) // 2
    for i in (0, 333, M - 1):
        exp = O.log_likelihood(O.OracleLocus(S.to_alignment_dict(aln[i])), _otree(new_trees[i]), om)
        assert rel_err(eng._felsen[i], exp) <= REL
    eng.restore_caches(snap)
    assert eng.log_likelihood(trees, None, 0.0) == total
    assert t2 != total


def test_program_template_same_topology_new_lengths():
    """All-dirty evaluations of an unchanged topology reuse the compiled op program and only refresh
    branch lengths (host fast path).  Results must equal a cold evaluation bit for bit, the cache
    left behind must support a correct dirty-path step, and NOSTORE scoring must agree."""
    P = _pkg()
    from phynetpy_b200 import simulate as S
    rng = np.random.default_rng(314)
    n = 14
    par, bl = S.random_trees(1, n, rng, height=0.5)
    aln = S.simulate_alignments(par, bl, 333, S.DEFAULT_PI, S.DEFAULT_RATES, rng)[0]
    d = S.to_alignment_dict(aln)
    ft = S.flat_trees(par, bl)[0]
    model = P.GTR(S.DEFAULT_PI, S.DEFAULT_RATES)
    om = O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES)
    locus = O.OracleLocus(d)
    fc = P.FelsensteinCalculator(d)
    ctx = fc._context()
    for k, scale in enumerate([1.0, 0.25, 0.5, 2.0, 4.0, 1.0]):   # the coupled moves' height scales
        t = ft.with_lengths(ft.blen * scale)
        got = fc.log_likelihood(t, model, full=True)
        assert ctx.last_stats()["nodes_recomputed"] == n - 1
        cold = P.FelsensteinCalculator(d)
        assert cold.log_likelihood(t, model) == got
        cold.close()
        assert rel_err(got, O.log_likelihood(locus, _otree(t), om)) <= REL
    # the cache image restored from the template describes the LAST tree: a one-branch change is a short path
    b2 = ft.blen.copy()
    b2[3] *= 1.3
    t2 = ft.with_lengths(b2)
    got = fc.log_likelihood(t2, model)
    assert 0 < ctx.last_stats()["nodes_recomputed"] < n - 1
    assert rel_err(got, O.log_likelihood(locus, _otree(t2), om)) <= REL
    # NOSTORE template: candidates that rescale every height
    eng = P.SeqLikelihoodEngine([d], None, model)
    eng.log_likelihood([ft], None, 0.0)
    cands = [ft.with_lengths(ft.blen * s) for s in (0.25, 0.5, 2.0, 4.0, 0.25)]
    vals = eng.score_candidates(0, cands)
    for t, v in zip(cands, vals):
        assert rel_err(v, O.log_likelihood(locus, _otree(t), om)) <= REL
    assert vals[0] == vals[4]
    # a polytomy with two absent taxa below one node (equal child refs): still correct
    gt = P.GeneTree.from_newick("((t0:0.1,zz1:0.2,zz2:0.05):0.1,(t1:0.1,t2:0.3):0.2,t3:0.2);")
    for scale in (1.0, 2.0, 0.5):
        t = gt.flat()
        t = t.with_lengths(t.blen * scale)
        got = fc.log_likelihood(t, model, full=True)
        assert rel_err(got, O.log_likelihood(locus, _otree(t), om)) <= REL


def test_large_tree_program_streamed_in_segments():
    """Trees whose op program exceeds one shared-memory segment (SEG_MAX_OPS = 192 ops, ~97 taxa):
    320 taxa -> 638 ops -> 4 segments.  Store mode, dirty path and NOSTORE (deep stash) vs the oracle."""
    P = _pkg()
    from phynetpy_b200 import simulate as S
    rng = np.random.default_rng(2718)
    n, L = 320, 150
    par, bl = S.random_trees(1, n, rng, height=0.25)
    aln = S.simulate_alignments(par, bl, L, S.DEFAULT_PI, S.DEFAULT_RATES, rng)[0]
    ft = S.flat_trees(par, bl)[0]
    taxa = S.tip_labels(n)
    model = P.GTR(S.DEFAULT_PI, S.DEFAULT_RATES)
    om = O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES)
    fc = P.FelsensteinCalculator.from_matrix(taxa, aln, device_compress=False)
    locus = O.OracleLocus(S.to_alignment_dict(aln))
    assert fc.n_patterns == locus.n_patterns
    exp = O.log_likelihood(locus, _otree(ft), om)
    assert exp > -600.0 * L  # not the clamp regime: a real comparison
    got = fc.log_likelihood(ft, model)
    assert rel_err(got, exp) <= REL
    assert fc._context().last_stats()["ops"] == 2 * (n - 1)
    b2 = ft.blen.copy()
    b2[11] *= 2.0
    t2 = ft.with_lengths(b2)
    got2 = fc.log_likelihood(t2, model)
    assert fc._context().last_stats()["nodes_recomputed"] < n - 1
    assert rel_err(got2, O.log_likelihood(locus, _otree(t2), om)) <= REL
    eng = P.SeqLikelihoodEngine([fc], None, model)
    vals = eng.score_candidates(0, [ft.with_lengths(ft.blen * 1.1), t2])
    assert rel_err(vals[0], O.log_likelihood(locus, _otree(ft.with_lengths(ft.blen * 1.1)), om)) <= REL
    assert vals[1] == got2


def test_many_candidates_one_locus_parallel_compile_with_templates():
    """The coupled moves' batch: one locus, hundreds of candidate trees ({g} U NNI(g)) x height scales,
    compiled by several host workers that share the locus' program template.  Every value must equal
    the single-tree evaluation bit for bit and the oracle to 1e-10."""
    P = _pkg()
    from phynetpy_b200 import simulate as S
    rng = np.random.default_rng(99)
    n = 16
    par, bl = S.random_trees(3, n, rng, height=0.4)
    aln = S.simulate_alignments(par[:1], bl[:1], 290, S.DEFAULT_PI, S.DEFAULT_RATES, rng)[0]
    d = S.to_alignment_dict(aln)
    topo = S.flat_trees(par, bl)            # three different topologies over the same taxa
    model = P.GTR(S.DEFAULT_PI, S.DEFAULT_RATES)
    om = O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES)
    locus = O.OracleLocus(d)
    eng = P.SeqLikelihoodEngine([d], None, model)
    eng.log_likelihood([topo[0]], None, 0.0)
    scales = (0.25, 0.5, 1.0, 2.0, 4.0)
    cands = []
    for k in range(300):                    # 300 x 31 nodes = 9300 nodes -> multi-threaded compile
        t = topo[int(rng.integers(0, 3))]
        cands.append(t.with_lengths(t.blen * scales[k % 5] * (1.0 + 0.001 * (k // 5))))
    vals = eng.score_candidates(0, cands)
    assert eng.last_stats["launches"] >= 2 and eng.last_stats["launches"] % 2 == 0  # (K-a, prune) per chunk
    fc = P.FelsensteinCalculator(d)
    for k in list(range(0, 300, 23)) + [299]:
        assert vals[k] == fc.log_likelihood(cands[k], model, full=True)
        assert rel_err(vals[k], O.log_likelihood(locus, _otree(cands[k]), om)) <= REL
    again = eng.score_candidates(0, cands)
    assert np.array_equal(vals, again)


def test_engine_entries_invalidated_behind_its_back():
    """INTEGRATION.md shares the scalar cache list with the reference engine: an entry set to None
    without invalidate_locus() must still be recomputed."""
    P = _pkg()
    ec = load_golden("engine_case.json")
    model = _model(P, ec["model"])
    eng = P.SeqLikelihoodEngine([dict(l["alignment"]) for l in ec["loci"]], None, model)
    trees = [_tree(P, l["tree"]) for l in ec["loci"]]
    total = eng.log_likelihood(trees, None, 0.0)
    eng._felsen[4] = None
    eng._felsen[7] = None
    assert eng.log_likelihood(trees, None, 0.0) == total
    assert all(v is not None for v in eng._felsen)


def test_out_device_on_caller_stream_matches_host_output():
    """PNB_EVAL_OUT_DEVICE + pnb_ctx_set_stream: the kernel writes per-locus logL straight into a
    caller-owned device vector (the NCCL buffer of the multi-GPU path) on the caller's stream."""
    import torch
    P = _pkg()
    from phynetpy_b200 import _lib
    from phynetpy_b200._seq_likelihood import _model_handle
    from phynetpy_b200.device import pack_trees
    ec = load_golden("engine_case.json")
    model = _model(P, ec["model"])
    calcs = [P.FelsensteinCalculator(dict(l["alignment"])) for l in ec["loci"]]
    trees = [_tree(P, l["tree"]).flat() for l in ec["loci"]]
    ctx = calcs[0]._context()
    for c in calcs:
        c._locus_handle()
    handles = np.array([c._locus.value for c in calcs], dtype=np.uint64)
    node_off, parent, blen, tip_row = pack_trees(trees, [c._row_of for c in calcs])
    mh = _model_handle(model, ctx)
    host = ctx.eval(mh, handles, node_off, parent, blen, tip_row, 1.0, _lib.EVAL_FULL)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    try:
        ctx.set_profiling(True)
        out = torch.full((len(calcs) + 2,), -1.0, dtype=torch.float64, device="cuda")
        with torch.cuda.stream(stream):
            ctx.eval(mh, handles, node_off, parent, blen, tip_row, 1.0, _lib.EVAL_FULL, None,
                     out_device_ptr=out.data_ptr() + 8)
            doubled = out * 2.0  # ordered after the kernel on the same stream
        stream.synchronize()
        tm = ctx.last_timings()
        assert tm["prune_ms"] > 0.0 and tm["call_ms"] > 0.0
        got = out.cpu().numpy()
        assert got[0] == -1.0 and got[-1] == -1.0
        assert np.array_equal(got[1:-1], host)
        assert np.array_equal(doubled.cpu().numpy()[1:-1], 2.0 * host)
        for a, b in zip(host, ec["per_locus"]):
            assert rel_err(a, b) <= REL
    finally:
        ctx.set_profiling(False)
        ctx.set_stream(None)


def test_rescale_mode_beyond_the_reference():
    """PNB_EVAL_RESCALE: (1) on trees where nothing underflows it agrees with the default mode (and
    hence the reference) to ~1e-13; (2) on a 1,200-taxon tree the default mode reproduces the
    reference's floor (every site at log(1e-300)), while the rescaled mode gives the true likelihood,
    checked against an independently rescaled NumPy pruning; (3) dirty path and NOSTORE work in it."""
    P = _pkg()
    from phynetpy_b200 import simulate as S
    model = P.GTR(S.DEFAULT_PI, S.DEFAULT_RATES)
    om = O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES)
    for case in CASES[:12]:
        fc = P.FelsensteinCalculator(dict(case["alignment"]))
        a = fc.log_likelihood(_tree(P, case["tree"]), _model(P, case["model"]), site_rate=case["site_rate"])
        b = fc.log_likelihood(_tree(P, case["tree"]), _model(P, case["model"]), site_rate=case["site_rate"], rescale=True)
        assert rel_err(b, a) <= 1e-12 and rel_err(b, case["logL"]) <= REL
    rng = np.random.default_rng(5150)
    n, L = 1200, 60
    par, bl = S.random_trees(1, n, rng, height=1.5)
    aln = S.simulate_alignments(par, bl, L, S.DEFAULT_PI, S.DEFAULT_RATES, rng)[0]
    ft = S.flat_trees(par, bl)[0]
    taxa = S.tip_labels(n)
    fc = P.FelsensteinCalculator.from_matrix(taxa, aln, device_compress=False)
    locus = O.OracleLocus(S.to_alignment_dict(aln))
    ref_like = O.log_likelihood(locus, _otree(ft), om)          # what the reference returns: the floor
    assert ref_like == pytest.approx(L * math.log(1e-300), rel=1e-12)
    assert rel_err(fc.log_likelihood(ft, model), ref_like) <= REL
    true_ll = O.log_likelihood_rescaled(locus, _otree(ft), om)
    assert true_ll < ref_like                                     # far below the floor
    got = fc.log_likelihood(ft, model, rescale=True)
    assert rel_err(got, true_ll) <= REL
    b2 = ft.blen.copy()
    b2[17] *= 3.0
    t2 = ft.with_lengths(b2)
    got2 = fc.log_likelihood(t2, model, rescale=True)
    assert fc._context().last_stats()["nodes_recomputed"] < 200
    assert rel_err(got2, O.log_likelihood_rescaled(locus, _otree(t2), om)) <= REL
    eng = P.SeqLikelihoodEngine([fc], None, model, rescale=True)
    assert eng.log_likelihood([t2], None, 0.0) == got2
    vals = eng.score_candidates(0, [ft, t2])
    assert vals[1] == got2 and rel_err(vals[0], true_ll) <= REL


def test_cfg5_shape_batch_split_invariance():
    """BASELINE cfg5 shape (32 taxa x 2,000 sites per locus) at 192 loci: the per-locus values do not
    depend on how the loci are batched (one call, two calls, locus by locus) -- the property that makes
    the multi-GPU sum independent of the number of ranks -- and a sample agrees with the oracle."""
    P = _pkg()
    from phynetpy_b200 import simulate as S
    rng = np.random.default_rng(55)
    M, n, L = 192, 32, 2000
    par, bl = S.random_trees(M, n, rng)
    aln = S.simulate_alignments(par, bl, L, S.DEFAULT_PI, S.DEFAULT_RATES, rng)
    trees = S.flat_trees(par, bl)
    taxa = S.tip_labels(n)
    model = P.GTR(S.DEFAULT_PI, S.DEFAULT_RATES)
    calcs = [P.FelsensteinCalculator.from_matrix(taxa, aln[i], device_compress=False) for i in range(M)]
    one = P.SeqLikelihoodEngine(calcs, None, model)
    total = one.log_likelihood(trees, None, 0.0)
    whole = list(one._felsen)
    one._commit_pending()   # the engines below share the loci: nothing may stay pending
    a = P.SeqLikelihoodEngine(calcs[:100], None, model)
    b = P.SeqLikelihoodEngine(calcs[100:], None, model)
    a.invalidate_device_cache(); b.invalidate_device_cache()
    ta = a.log_likelihood(trees[:100], None, 0.0)
    tb = b.log_likelihood(trees[100:], None, 0.0)
    a._commit_pending(); b._commit_pending()
    assert a._felsen + b._felsen == whole
    assert rel_err(ta + tb, total) <= 1e-14
    for i in (0, 77, 191):
        assert calcs[i].log_likelihood(trees[i], model, full=True) == whole[i]
    om = O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES)
    for i in (3, 150):
        exp = O.log_likelihood(O.OracleLocus(S.to_alignment_dict(aln[i])), _otree(trees[i]), om)
        assert rel_err(whole[i], exp) <= REL


def test_several_engines_in_one_device_call():
    """MC3 / several chains on one device: the dirty loci of all engines go out in ONE pnb_eval; values, pending
    state, commit and rollback behave exactly as for engines evaluated one by one (reference MC3 loop
    src/_mcmc_seq.py:3584-3723 evaluates its chains one after the other)."""
    from phynetpy_b200 import simulate as S
    P = _pkg()
    rng = np.random.default_rng(31)
    M, n, L, T = 12, 9, 150, 3
    par, bl = S.random_trees(M, n, rng)
    aln = S.simulate_alignments(par, bl, L, S.DEFAULT_PI, S.DEFAULT_RATES, rng)
    loci = [S.to_alignment_dict(aln[i]) for i in range(M)]
    model = P.GTR(S.DEFAULT_PI, S.DEFAULT_RATES)
    chains = [P.SeqLikelihoodEngine(loci, None, model) for _ in range(T)]
    solo = [P.SeqLikelihoodEngine(loci, None, model) for _ in range(T)]
    # every chain has its own gene trees
    trees = []
    for c in range(T):
        p2, b2 = S.random_trees(M, n, rng)
        trees.append(S.flat_trees(p2, b2))
    many = P.SeqLikelihoodEngine.log_likelihood_many(chains, trees)
    assert chains[0].last_stats["launches"] <= 2                     # one evaluation for 3 x 12 loci
    one = [solo[c].log_likelihood(trees[c], None, 0.0) for c in range(T)]
    assert many == one
    assert [e._felsen for e in chains] == [e._felsen for e in solo]
    # one gene-tree move per chain, different loci: accept in chain 0 and 2, reject in chain 1
    new = [list(t) for t in trees]
    moved = [2, 7, 11]
    for c in range(T):
        b = trees[c][moved[c]].blen.copy()
        b[0] *= 1.5
        new[c][moved[c]] = trees[c][moved[c]].with_lengths(b)
    snaps = [e.clone_caches() for e in chains]
    ssnaps = [e.clone_caches() for e in solo]
    for c in range(T):
        chains[c].invalidate_locus(moved[c])
        solo[c].invalidate_locus(moved[c])
    many2 = P.SeqLikelihoodEngine.log_likelihood_many(chains, new)
    one2 = [solo[c].log_likelihood(new[c], None, 0.0) for c in range(T)]
    assert many2 == one2 and many2 != many
    assert chains[0].last_stats["nodes_recomputed"] <= 3 * (n - 1)      # dirty paths only, all chains together
    chains[1].restore_caches(snaps[1])
    solo[1].restore_caches(ssnaps[1])
    cur = [new[0], trees[1], new[2]]
    many3 = P.SeqLikelihoodEngine.log_likelihood_many(chains, cur)
    one3 = [solo[c].log_likelihood(cur[c], None, 0.0) for c in range(T)]
    assert many3 == one3 and many3[1] == many[1] and many3[0] == many2[0]
    # the rejected proposal left no trace on the device: re-proposing it recomputes the same value
    chains[1].invalidate_locus(moved[1])
    again = P.SeqLikelihoodEngine.log_likelihood_many(chains, new)
    assert again == many2
    # a full recomputation of every chain agrees bit for bit
    for c in range(T):
        for i in range(M):
            chains[c].invalidate_locus(i)
        chains[c].invalidate_device_cache()
    assert P.SeqLikelihoodEngine.log_likelihood_many(chains, new) == many2
    with pytest.raises(ValueError, match="one gene-tree list per engine"):
        P.SeqLikelihoodEngine.log_likelihood_many(chains, new[:2])
    other = P.SeqLikelihoodEngine(loci, None, P.JC69())
    with pytest.raises(ValueError, match="share context, model"):
        P.SeqLikelihoodEngine.log_likelihood_many([chains[0], other], [new[0], new[0]])


def test_resident_program_is_not_rewritten_under_a_batch_that_reads_it():
    """Score-only batches on a locus WITHOUT a committed cache: batch [Z] makes Z's program device-resident;
    batch [Z, A, Z, B, Z] lets a later job of the same locus rebuild the template (other topology, same op count)
    while an earlier job of that batch reads the resident copy in place.  Every value must still be its own
    tree's (round-1 advisory: the rebuilt program used to be copied over the buffer the earlier job read)."""
    P = _pkg()
    from phynetpy_b200 import simulate as S
    rng = np.random.default_rng(2024)
    n = 12
    par, bl = S.random_trees(3, n, rng, height=0.3)
    aln = S.simulate_alignments(par[:1], bl[:1], 400, S.DEFAULT_PI, S.DEFAULT_RATES, rng)[0]
    d = S.to_alignment_dict(aln)
    Z, A, B = S.flat_trees(par, bl)
    model = P.GTR(S.DEFAULT_PI, S.DEFAULT_RATES)
    om = O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES)
    locus = O.OracleLocus(d)
    want = {id(t): O.log_likelihood(locus, _otree(t), om) for t in (Z, A, B)}
    for rep in range(3):
        eng = P.SeqLikelihoodEngine([d], None, model)        # fresh, uncommitted locus each time
        v1 = eng.score_candidates(0, [Z])
        assert rel_err(v1[0], want[id(Z)]) <= REL
        batch = [Z, A, Z, B, Z, A]
        v2 = eng.score_candidates(0, batch)
        for t, v in zip(batch, v2):
            assert rel_err(v, want[id(t)]) <= REL, (rep, v, want[id(t)])
        assert v2[0] == v2[2] == v2[4] == v1[0] and v2[1] == v2[5]
        # many repetitions in one batch: several workers, several rebuilds
        big = [Z, A, B] * 200
        v3 = eng.score_candidates(0, big)
        assert all(v3[k] == v2[(0, 1, 3)[k % 3]] for k in range(len(big)))


def test_split_tail_same_bits(monkeypatch):
    """PNB_SPLIT_TAIL=1 runs the tiles of a launch's last, partial wave as two half-tile blocks each (every launch of at
    most half a wave is all half tiles).  A tile's two half sums are formed the same way by one block or by two, so the
    log-likelihoods are the same bits: single loci of 1 .. 5 tiles (full and dirty-path evaluations, stored, score-only
    and rescaled) and a batch of 700 small loci (more than one wave of blocks, the tail split)."""
    P = _pkg()
    from phynetpy_b200 import simulate as S
    rng = np.random.default_rng(314)
    model = P.GTR(S.DEFAULT_PI, S.DEFAULT_RATES)

    def run():
        out = []
        r = np.random.default_rng(99)
        for n, L in ((7, 300), (12, 1400), (20, 2600)):
            par, bl = S.random_trees(1, n, r)
            aln = S.simulate_alignments(par, bl, L, S.DEFAULT_PI, S.DEFAULT_RATES, r)[0]
            ft = S.flat_trees(par, bl)[0]
            for rescale in (False, True):
                fc = P.FelsensteinCalculator(S.to_alignment_dict(aln))
                out.append(fc.log_likelihood(ft, model, rescale=rescale))
                bl2 = ft.blen.copy()
                bl2[1] *= 1.3
                out.append(fc.log_likelihood(ft.with_lengths(bl2), model, rescale=rescale))          # dirty path
                fc.close()
        M = 700
        par, bl = S.random_trees(M, 8, r)
        aln = S.simulate_alignments(par, bl, 200, S.DEFAULT_PI, S.DEFAULT_RATES, r)
        eng = P.SeqLikelihoodEngine([S.to_alignment_dict(aln[i]) for i in range(M)], None, model)
        trees = S.flat_trees(par, bl)
        out.append(eng.log_likelihood(trees, None, 0.0))
        out.extend(eng._felsen)
        return out

    monkeypatch.setenv("PNB_SPLIT_TAIL", "0")
    a = run()
    monkeypatch.setenv("PNB_SPLIT_TAIL", "1")
    b = run()
    monkeypatch.delenv("PNB_SPLIT_TAIL", raising=False)
    assert a == b
