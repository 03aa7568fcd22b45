"""Candidate sets of the coupled moves: enumeration on flat arrays (CPU, bit-exact signatures against
the reference's ``_candidate_set``) and their scores in one device call (GPU, 1e-10)."""
import numpy as np
import pytest

from phynetpy_b200.candidates import (HEIGHT_SCALE_GRID, CandidateBatch, _candidate_arrays_plain, candidate_arrays,
                                      candidate_set, candidate_table, find_member_in_batch, nni_neighbors, node_heights,
                                      signature)
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


def _odd_trees():
    # node order that is not children-first, a polytomy, all-zero heights (scales collide after rounding)
    return [FlatTree([4, 4, 3, -1, 3], [0.1, 0.1, 0.3, None, 0.2], ["a", "b", "c", None, None]),
            FlatTree([3, 3, 3, 5, 5, -1], [0.1, 0.1, 0.1, 0.2, 0.3, None], ["a", "b", "c", None, "d", None]),
            FlatTree([2, 2, 4, 4, -1], [0.0, 0.0, 0.0, 0.0, None], ["a", "b", None, "c", None])]


@pytest.mark.parametrize("scales", [HEIGHT_SCALE_GRID, (0.3, 1.0, 1.7), (1.0,), (2.0, 0.5)])
def test_native_enumeration_equals_python(scales):
    """pnb_candidates_count / pnb_candidates_fill (host code behind the C ABI, all trees in one call) against the
    NumPy enumeration and the plain one: members, order, every bit -- for any scale grid."""
    from phynetpy_b200 import simulate as S
    rng = np.random.default_rng(21)
    trees = _odd_trees()
    for n in (2, 3, 5, 8, 16, 33, 64):
        par, bl = S.random_trees(4, n, rng)
        trees += S.flat_trees(par, bl)
    hs = [node_heights(t) for t in trees]
    cb = CandidateBatch(trees, hs, scales)
    assert cb.member_off[0] == 0 and cb.node_off[-1] == cb.parent.shape[0] == cb.blen.shape[0] == cb.height.shape[0]
    for j, (t, h) in enumerate(zip(trees, hs)):
        P, L, H = candidate_table(t, h, scales)
        p, l, hh = cb.table(j)
        assert p.shape == P.shape
        assert np.array_equal(p, P) and np.array_equal(hh, H)
        assert np.array_equal(np.nan_to_num(l, nan=-1.0), np.nan_to_num(L, nan=-1.0))
        if t.n_nodes <= 31:
            a = _candidate_arrays_plain(t, scales)
            assert np.array_equal(a[0], p) and np.array_equal(np.nan_to_num(a[1], nan=-1.0), np.nan_to_num(l, nan=-1.0))
        m0 = int(cb.member_off[j])
        assert (cb.tree_of_member[m0:int(cb.member_off[j + 1])] == j).all()
        assert cb.changed_node[m0] == -1                                  # the tree itself comes first
        for k in range(0, p.shape[0], 5):
            assert find_member_in_batch(cb, j, p[k], hh[k]) == k
    # expand(): one per-node array per tree, repeated for every member of that tree
    per_tree = [np.arange(t.n_nodes, dtype=np.int32) + 1000 * j for j, t in enumerate(trees)]
    ex = cb.expand(per_tree, np.int32)
    assert ex.shape[0] == cb.parent.shape[0]
    for m in range(0, cb.n_members, 37):
        a, b = int(cb.node_off[m]), int(cb.node_off[m + 1])
        assert np.array_equal(ex[a:b], per_tree[int(cb.tree_of_member[m])])


def test_native_enumeration_rejects_malformed_trees():
    with pytest.raises(ValueError, match="out of range"):
        CandidateBatch([FlatTree([5, 2, -1], [0.1, 0.1, None], ["a", "b", None])], [np.zeros(3)])
    with pytest.raises(ValueError, match="no root|cycle"):
        CandidateBatch([FlatTree([1, 0], [0.1, 0.1], ["a", None])], [np.zeros(2)])
    with pytest.raises(ValueError, match="more than 128"):
        from phynetpy_b200 import simulate as S
        par, bl = S.random_trees(1, 70, np.random.default_rng(0))              # 139 nodes
        big = S.flat_trees(par, bl)[0]
        CandidateBatch([big], [node_heights(big)])
    with pytest.raises(ValueError, match="one height per node"):
        CandidateBatch([FlatTree([1, -1], [0.1, None], ["a", None])], [np.zeros(3)])
    assert CandidateBatch([], []).n_members == 0
