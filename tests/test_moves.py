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


# ---------------------------------------------------------------------------------------------------
# species-network parameter moves
# ---------------------------------------------------------------------------------------------------
NET_CASES = json.load(open(os.path.join(HERE, "golden", "moves_cases.json")))["network_cases"]


@pytest.mark.parametrize("case", NET_CASES, ids=[c["name"] for c in NET_CASES])
def test_network_height_slide_reproduces_the_reference(case):
    from phynetpy_b200._msnc_density import SpeciesNetwork
    net = SpeciesNetwork.from_dict(case["network"])
    label = net.node_label
    cands = MV.net_slide_candidates(net)
    for rec in case["slides"]:
        assert rec["n_candidates"] == cands.shape[0]
        node = label.index(rec["node"])
        assert node in cands.tolist()
        out = MV.net_slide_node(net, node, rec["u"])
        if rec["result"] is None:
            assert out is None
            continue
        new, loghr = out
        assert new is not net and float(new.height[node]) == rec["result"]["height"] and loghr == rec["result"]["loghr"]
        assert np.array_equal(np.delete(new.height, node), np.delete(net.height, node))
        assert np.array_equal(new.edge_gamma, net.edge_gamma, equal_nan=True) and new.species == net.species
        # a reticulation never rises above its LOWER parent
        parents = net.edge_src[net.edge_dst == node]
        if parents.shape[0] == 2:
            assert new.height[node] < net.height[parents].min()


@pytest.mark.parametrize("case", NET_CASES, ids=[c["name"] for c in NET_CASES])
def test_gamma_move_reproduces_the_reference(case):
    from phynetpy_b200._msnc_density import SpeciesNetwork
    net = SpeciesNetwork.from_dict(case["network"])
    label = net.node_label
    rets = MV.reticulations(net)
    for rec in case["gammas"]:
        if rec["result"] is None:
            assert rets.shape[0] == 0 and MV.gamma_move(net, np.random.default_rng(rec["seed"])) is None
            continue
        assert rec["n_reticulations"] == rets.shape[0]
        r = label.index(rec["result"]["retic"])
        edges = np.nonzero(net.edge_dst == r)[0]
        by_parent = {label[int(net.edge_src[e])]: int(e) for e in edges}
        assert {p: float(net.edge_gamma[e]) for p, e in by_parent.items()} == rec["result"]["old"]
        # the reference steps the gamma of whichever in-edge its (unordered) edge set lists first: try both
        outcomes = []
        for first in edges.tolist():
            new = MV.gamma_move_at(net, r, rec["u"], rec["window"], first_edge=first)
            outcomes.append({p: float(new.edge_gamma[e]) for p, e in by_parent.items()})
            others = np.ones(net.edge_gamma.shape[0], bool)
            others[edges] = False
            assert np.array_equal(new.edge_gamma[others], net.edge_gamma[others], equal_nan=True)
            assert sum(outcomes[-1].values()) == pytest.approx(1.0, abs=1e-15) and all(0 < g < 1 for g in outcomes[-1].values())
        assert rec["result"]["new"] in outcomes


def test_network_moves_feed_the_oracle_density_consistently():
    """A moved network is a valid network: the density of a gene tree changes continuously with a small slide and
    the old object is untouched."""
    from oracle import msnc_oracle as MO
    from phynetpy_b200 import simulate as S
    rng = np.random.default_rng(9)
    net = S.random_species_network(6, 2, rng)
    trees, species_of = S.simulate_gene_trees_msnc(net, 1, 0.02, rng, 3)
    h0, g0 = net.height.copy(), net.edge_gamma.copy()
    base = [MO.msnc_log_density(net.to_dict(), t.parent.tolist(), t.blen.tolist(), t.label, species_of, 0.02) for t in trees]
    n_moved = 0
    for step in range(40):
        out = MV.net_slide_height_move(net, rng) if step % 2 else MV.gamma_move(net, rng)
        if out is None:
            continue
        new, loghr = out
        assert math.isfinite(loghr) and new is not net
        assert (new.height[new.edge_src] >= new.height[new.edge_dst]).all()           # still time-consistent
        d = [MO.msnc_log_density(new.to_dict(), t.parent.tolist(), t.blen.tolist(), t.label, species_of, 0.02) for t in trees]
        assert all(math.isfinite(x) or x == -math.inf for x in d)
        n_moved += 1
    assert n_moved > 20
    assert np.array_equal(net.height, h0) and np.array_equal(net.edge_gamma, g0, equal_nan=True)
    th, lh = MV.theta_move(0.02, np.random.default_rng(1))
    assert th == 0.02 * math.exp(lh) and abs(lh) <= 0.2
