"""Candidate sets of the coupled moves: enumeration on flat arrays (CPU, bit-exact signatures against
the reference's ``_candidate_set``) and their scores in one device call (GPU, 1e-10)."""
import numpy as np
import pytest

from phynetpy_b200.candidates import (HEIGHT_SCALE_GRID, _candidate_arrays_plain, candidate_arrays, candidate_set,
                                      candidate_table, nni_neighbors, node_heights, signature)
from phynetpy_b200.tree import FlatTree
from conftest import load_golden

CASES = load_golden("candidate_cases.json")["cases"]


def _sig_key(sig):
    return tuple(sorted((tuple(sorted(ls)), h) for ls, h in sig))


def _golden_key(entry):
    return tuple(sorted((tuple(ls), h) for ls, h in entry["signature"]))


@pytest.mark.parametrize("case", CASES, ids=["n%d" % ((len(c["tree"]["parent"]) + 1) // 2) for c in CASES])
def test_candidate_set_equals_reference(case):
    t = case["tree"]
    ft = FlatTree(t["parent"], t["blen"], t["label"])
    cands = candidate_set(ft)
    assert len(cands) == case["n_candidates"]
    ours = {_sig_key(signature(c)) for c in cands}
    ref = {_golden_key(e) for e in case["candidates"]}
    assert ours == ref                                  # same topologies, same rounded heights
    assert _sig_key(signature(ft)) in ours              # the identity is always a member
    ident = [c for c in cands if _sig_key(signature(c)) == _sig_key(signature(ft))]
    assert len(ident) == 1 and np.array_equal(ident[0].parent, ft.parent)
    assert np.array_equal(ident[0].blen, ft.blen, equal_nan=True)   # untouched at scale 1, as in the reference
    # height-preserving NNI: node heights unchanged, lengths re-derived from them
    h = node_heights(ft)
    for nb, hh in nni_neighbors(ft):
        assert np.array_equal(hh, h)
        np.testing.assert_allclose(node_heights(nb), h, rtol=0, atol=1e-15)
    # the stacked-array form holds the same trees in the same order
    par, bl = candidate_arrays(ft)
    assert par.shape == (len(cands), ft.n_nodes)
    for k, c in enumerate(cands):
        assert np.array_equal(par[k], c.parent) and np.array_equal(bl[k], c.blen, equal_nan=True)
    # the scale grid is closed under reciprocals: scaling a candidate back reproduces its heights exactly
    assert all(1.0 / s in HEIGHT_SCALE_GRID for s in HEIGHT_SCALE_GRID)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=["n%d" % ((len(c["tree"]["parent"]) + 1) // 2) for c in CASES])
def test_candidate_scores_match_reference(case):
    from phynetpy_b200 import infer as P
    t = case["tree"]
    ft = FlatTree(t["parent"], t["blen"], t["label"])
    model = P.GTR(case["model"]["pi"], case["model"]["rates"])
    eng = P.SeqLikelihoodEngine([dict(case["alignment"])], None, model)
    eng.log_likelihood([ft], None, 0.0)
    cands = candidate_set(ft)
    (vals,) = eng.score_candidate_sets([(0, cands)])
    ref = {_golden_key(e): e["logL"] for e in case["candidates"]}
    for c, v in zip(cands, vals):
        exp = ref[_sig_key(signature(c))]
        assert abs(v - exp) <= 1e-10 * abs(exp)
    # all loci x all candidates in ONE call (here: the same locus three times)
    res = eng.score_candidate_sets([(0, cands), (0, cands[:3]), (0, cands)])
    assert np.array_equal(res[0], vals) and np.array_equal(res[2], vals) and np.array_equal(res[1], vals[:3])
    par, bl = candidate_arrays(ft)
    (vals2,) = eng.score_candidate_sets([(0, (ft, par, bl))])
    assert np.array_equal(vals2, vals)


def test_fast_enumeration_equals_the_plain_one():
    """candidate_table tells NNI neighbours apart by (node, new clade) and scales lengths by exact powers of two
    instead of building a tree and hashing a signature per member; members, order and every bit must agree."""
    from phynetpy_b200 import simulate as S
    rng = np.random.default_rng(8)
    for n in (2, 3, 4, 6, 9, 16, 25):
        par, bl = S.random_trees(12, n, rng)
        for ft in S.flat_trees(par, bl):
            a = _candidate_arrays_plain(ft)
            p, l, h = candidate_table(ft)
            assert np.array_equal(a[0], p)
            assert np.array_equal(np.nan_to_num(a[1], nan=-1.0), np.nan_to_num(l, nan=-1.0))
            for k in range(0, p.shape[0], 7):
                assert np.allclose(node_heights(FlatTree(p[k], l[k], ft.label)), h[k], rtol=0, atol=1e-12)
    # node order that is not children-first, a polytomy, all-zero heights (scales collide after rounding)
    for ft in (FlatTree([4, 4, 3, -1, 3], [0.1, 0.1, 0.3, None, 0.2], ["a", "b", "c", None, None]),
               FlatTree([3, 3, 3, 5, 5, -1], [0.1, 0.1, 0.1, 0.2, 0.3, None], ["a", "b", "c", None, "d", None]),
               FlatTree([2, 2, 4, 4, -1], [0.0, 0.0, 0.0, 0.0, None], ["a", "b", None, "c", None])):
        a = _candidate_arrays_plain(ft)
        p, l, h = candidate_table(ft)
        assert np.array_equal(a[0], p) and np.array_equal(np.nan_to_num(a[1], nan=-1.0), np.nan_to_num(l, nan=-1.0))
    assert np.array_equal(candidate_arrays(ft)[0], p)
