"""Gene-tree proposals on flat arrays (phynetpy_b200/moves.py) against golden vectors recorded from the
reference's operators (tests/golden/make_golden_moves.py).  Host logic only."""
import json
import math
import os

import numpy as np
import pytest

from phynetpy_b200 import moves as MV
from phynetpy_b200.candidates import node_heights
from phynetpy_b200.tree import FlatTree

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "moves_cases.json")))["cases"]


def _tree(case):
    return FlatTree(case["parent"], case["blen"], case["label"])


def _clades(ft):
    kids = [[] for _ in range(ft.n_nodes)]
    for i, p in enumerate(ft.parent.tolist()):
        if p >= 0:
            kids[p].append(i)
    out = [None] * ft.n_nodes

    def rec(v):
        if out[v] is None:
            out[v] = sorted(sum((rec(c) for c in kids[v]), [])) if kids[v] else [ft.label[v]]
        return out[v]
    for v in range(ft.n_nodes):
        rec(v)
    return out, kids


def _describe(ft, h):
    cl, kids = _clades(ft)
    internal = sorted(([cl[v], float(h[v])] for v in range(ft.n_nodes) if kids[v]), key=lambda x: (x[0], x[1]))
    lengths = sorted([cl[i], cl[p], float(ft.blen[i])] for i, p in enumerate(ft.parent.tolist()) if p >= 0)
    return {"internal": internal, "lengths": lengths}


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_slide_reproduces_the_reference_outcomes(case):
    ft = _tree(case)
    h = node_heights(ft)
    cl, _ = _clades(ft)
    cands = MV.slide_candidates(ft)
    root = int(np.nonzero(ft.parent < 0)[0][0])
    assert cands[-1] == root
    for rec in case["slides"]:
        assert rec["n_candidates"] == cands.shape[0]
        node = next(v for v in cands.tolist() if cl[v] == rec["node"] and (v == root) == rec["is_root"])
        out = MV.slide_node(ft, h, node, rec["u"])
        if rec["result"] is None:
            assert out is None
            continue
        h2, loghr = out
        assert h2[node] == rec["result"]["height"]            # same formula, same bits
        assert loghr == rec["result"]["loghr"]
        assert np.array_equal(np.delete(h2, node), np.delete(h, node))
        if "tree" in rec["result"]:
            moved = MV._rebuilt(ft, ft.parent, h2, True)
            assert _describe(moved, h2) == rec["result"]["tree"]     # _sync_lengths, to the bit


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_nni_outcome_set_equals_the_reference(case):
    ft = _tree(case)
    h = node_heights(ft)
    mine, n_invalid = {}, 0
    for p, v, s in MV.nni_eligible(ft):
        for c in np.nonzero(ft.parent == v)[0].tolist():
            t = MV.nni_swap(ft, h, v, c)
            if t is None:
                n_invalid += 1
                assert h[v] <= h[s]
                continue
            d = _describe(t, h)
            mine[json.dumps(d, sort_keys=True)] = d
            assert np.allclose(node_heights(t), h, rtol=0, atol=1e-15)      # height-preserving
    ref = {json.dumps(d, sort_keys=True) for d in case["nni_outcomes"]}
    assert set(mine) == ref
    assert (case["nni_rejected"] > 0) == (n_invalid > 0 or not MV.nni_eligible(ft))


def test_random_moves_keep_the_tree_valid_and_the_input_untouched():
    rng = np.random.default_rng(3)
    case = CASES[-1]
    ft = _tree(case)
    h = node_heights(ft)
    parent0, blen0, h0 = ft.parent.copy(), ft.blen.copy(), h.copy()
    cur, cur_h, moved, swapped = ft, h, 0, 0
    for step in range(400):
        if step % 2:
            out = MV.slide_height_move(cur, cur_h, rng)
            if out is None:
                continue
            nxt, nh, loghr = out
            assert math.isfinite(loghr)
            moved += 1
        else:
            out = MV.nni_move(cur, cur_h, rng)
            if out is None:
                continue
            nxt, nh = out
            swapped += 1
        assert nxt is not cur                                   # copy-and-swap: a fresh object per change
        assert (np.nan_to_num(nxt.blen) >= 0).all()
        assert np.allclose(node_heights(nxt), nh, rtol=0, atol=1e-12)
        assert sorted(l for l in nxt.label if l) == sorted(l for l in ft.label if l)
        assert (nxt.n_children()[np.array([l is not None for l in nxt.label])] == 0).all()
        cur, cur_h = nxt, nh
    assert moved > 50 and swapped > 50
    assert np.array_equal(ft.parent, parent0) and np.array_equal(np.nan_to_num(ft.blen), np.nan_to_num(blen0))
    assert np.array_equal(h, h0)


def test_draw_order_follows_the_reference():
    # node choice first, then the uniform (src/_mcmc_seq.py:969-989); one child draw per eligible node, then the tuple
    ft = _tree(CASES[2])
    h = node_heights(ft)
    r1, r2 = np.random.default_rng(11), np.random.default_rng(11)
    out = MV.slide_height_move(ft, h, r1)
    cands = MV.slide_candidates(ft)
    node = int(cands[int(r2.integers(cands.shape[0]))])
    exp = MV.slide_node(ft, h, node, float(r2.random()))
    assert (out is None) == (exp is None)
    if out is not None:
        assert np.array_equal(out[1], exp[0]) and out[2] == exp[1]
    assert r1.random() == r2.random()
    r1, r2 = np.random.default_rng(12), np.random.default_rng(12)
    MV.nni_move(ft, h, r1)
    el = MV.nni_eligible(ft)
    for p, v, s in el:
        r2.integers(int((ft.parent == v).sum()))
    r2.integers(len(el))
    assert r1.random() == r2.random()
