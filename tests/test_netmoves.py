"""Reticulation surgery and the placement scan of the informed add / delete moves on flat networks
(phynetpy_b200/netmoves.py) against the reference (tests/golden/netmoves_cases.json, prior patched to 0).  Host
logic; the densities come from the oracle here and from one ``pnb_msnc_eval`` per candidate network in production."""
import json
import math
import os

import numpy as np
import pytest

from oracle import msnc_oracle as MO
from phynetpy_b200 import netmoves as NM
from phynetpy_b200._msnc_density import SpeciesNetwork
from phynetpy_b200.candidates import HEIGHT_SCALE_GRID
from phynetpy_b200.tree import FlatTree

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "netmoves_cases.json")))["cases"]


def _labels(net, e):
    return [net.node_label[int(net.edge_src[e])], net.node_label[int(net.edge_dst[e])]]


def _canonical(net):
    return {"nodes": sorted([net.node_label[v], float(net.height[v])] for v in range(net.n_nodes)),
            "edges": sorted([net.node_label[int(s)], net.node_label[int(d)], None if g != g else float(g)]
                            for s, d, g in zip(net.edge_src, net.edge_dst, net.edge_gamma.tolist()))}


def _density_fn(case, scaled):
    def fn(cand):
        d = cand.to_dict()
        return np.array([MO.msnc_log_density(d, t.parent.tolist(), t.blen.tolist(), t.label, case["species_of"], case["theta"])
                         for t in scaled])
    return fn


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_geometric_placements_equal_the_reference(case):
    net = SpeciesNetwork.from_dict(case["network"])
    mine = sorted([_labels(net, d), _labels(net, r)] for d, r in NM.geometric_placements(net))
    assert mine == sorted(case["geometric"])


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_scored_placements_equal_the_reference(case):
    net = SpeciesNetwork.from_dict(case["network"])
    trees = [FlatTree(t["parent"], t["blen"], t["label"]) for t in case["trees"]]
    scaled = NM.scaled_gene_trees(trees)
    assert len(scaled) == len(trees) * len(HEIGHT_SCALE_GRID)
    fn = _density_fn(case, scaled)
    # the score of the network itself (reference _network_only_log_score with the prior at 0)
    assert NM.network_only_msnc_score(fn(net), len(HEIGHT_SCALE_GRID)) == pytest.approx(case["base_score"], rel=1e-12)
    got = NM.add_reticulation_placements(net, fn, len(HEIGHT_SCALE_GRID))
    mine = {(tuple(_labels(net, p["donor_edge"])), tuple(_labels(net, p["retic_edge"]))): p for p in got}
    ref = {(tuple(p["donor"]), tuple(p["retic"])): p for p in case["placements"]}
    assert set(mine) == set(ref)
    for k, p in ref.items():
        assert mine[k]["l2"] == pytest.approx(p["l2"], rel=1e-13)
        assert mine[k]["s_rep"] == pytest.approx(p["s_rep"], rel=1e-11)
    # the selection the add move makes from these scores
    s = np.array([p["s_rep"] for p in got])
    lse = s.max() + math.log(np.exp(s - s.max()).sum())
    assert abs(np.exp(s - lse).sum() - 1.0) < 1e-12


@pytest.mark.parametrize("case", [c for c in CASES if c["deletes"]], ids=[c["name"] for c in CASES if c["deletes"]])
def test_delete_surgery_equals_the_reference(case):
    net = SpeciesNetwork.from_dict(case["network"])
    lab = net.node_label
    for d in case["deletes"]:
        out = NM.remove_reticulation(net, lab.index(d["retic"]), lab.index(d["donor_parent"]))
        if d["result"] is None:
            assert out is None
            continue
        assert out is not None
        got = _canonical(out)
        assert got["nodes"] == [[a, pytest.approx(b, abs=1e-15)] for a, b in d["result"]["nodes"]]
        assert got["edges"] == d["result"]["edges"]
        assert out.n_reticulations() == net.n_reticulations() - 1 and NM.is_acyclic(out) and not NM.has_parallel_edges(out)


def test_insert_then_remove_is_the_identity():
    rng = np.random.default_rng(3)
    for case in CASES:
        net = SpeciesNetwork.from_dict(case["network"])
        base = _canonical(net)
        done = 0
        for d, r in NM.geometric_placements(net):
            rep = NM.representative_placement(net, d, r)
            if rep is None:
                continue
            cand, l2 = rep
            assert cand.n_nodes == net.n_nodes + 2 and cand.n_reticulations() == net.n_reticulations() + 1
            assert (cand.height[cand.edge_src] >= cand.height[cand.edge_dst]).all()
            v1, v2 = net.n_nodes, net.n_nodes + 1
            back = NM.remove_reticulation(cand, v2, v1)
            if back is None:
                continue
            got = _canonical(back)
            # gammas: a split drops the inheritance probability of the edge it splits (as the reference does), so
            # compare topology and heights, and the gammas wherever the original edge was not split
            assert got["nodes"] == base["nodes"]
            assert sorted(e[:2] for e in got["edges"]) == sorted(e[:2] for e in base["edges"])
            done += 1
        assert done >= 4
    assert NM.insert_reticulation(net, 0, 0, 0.1, 0.05, 0.5) is None


def test_device_density_fn_plumbing(monkeypatch):
    """device_density_fn hands the scaled trees and the candidate network to msnc_log_density_batch (one call per
    candidate); with the oracle behind that call the scan reproduces the reference."""
    case = CASES[1]
    calls = []

    def fake_batch(trees, net, species_of, theta, *, branch_thetas=None, context=None):
        calls.append(len(trees))
        d = net.to_dict()
        return np.array([MO.msnc_log_density(d, t.parent.tolist(), t.blen.tolist(), t.label, species_of, theta) for t in trees])
    monkeypatch.setattr(NM, "msnc_log_density_batch", fake_batch)
    net = SpeciesNetwork.from_dict(case["network"])
    trees = [FlatTree(t["parent"], t["blen"], t["label"]) for t in case["trees"]]
    fn = NM.device_density_fn(trees, case["species_of"], case["theta"])
    got = NM.add_reticulation_placements(net, fn, len(HEIGHT_SCALE_GRID), prior_fn=lambda n: -1.5)
    assert len(got) == len(case["placements"]) and set(calls) == {len(trees) * len(HEIGHT_SCALE_GRID)}
    ref = {(tuple(p["donor"]), tuple(p["retic"])): p["s_rep"] for p in case["placements"]}
    for p in got:
        k = (tuple(_labels(net, p["donor_edge"])), tuple(_labels(net, p["retic_edge"])))
        assert p["s_rep"] == pytest.approx(ref[k] - 1.5, rel=1e-11)          # the caller's prior is added to every score


def _keyed(net, old_labels):
    def key(v):
        lab = net.node_label[v]
        return lab if lab in old_labels else "new@%r" % float(net.height[v])
    return {"nodes": sorted([key(v), float(net.height[v])] for v in range(net.n_nodes)),
            "edges": sorted([key(int(s)), key(int(d)), None if g != g else float(g)]
                            for s, d, g in zip(net.edge_src, net.edge_dst, net.edge_gamma.tolist()))}


def _placements_fn(case):
    trees = [FlatTree(t["parent"], t["blen"], t["label"]) for t in case["trees"]]
    scaled = NM.scaled_gene_trees(trees)

    def fn(net):
        return NM.add_reticulation_placements(net, _density_fn(case, scaled), len(HEIGHT_SCALE_GRID))
    return fn


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_informed_add_move_equals_the_reference(case):
    """op_add_reticulation (prior at 0): for the placement and the three uniforms the reference drew, the same
    network (new nodes named by height) and the same log Hastings ratio."""
    net = SpeciesNetwork.from_dict(case["network"])
    old = set(net.node_label)
    placements = _placements_fn(case)(net)
    index = {(tuple(_labels(net, p["donor_edge"])), tuple(_labels(net, p["retic_edge"]))): k for k, p in enumerate(placements)}
    for rec in case["add_moves"]:
        k = index[(tuple(rec["donor"]), tuple(rec["retic"]))]
        if len(rec["uniforms"]) < 3:       # the reference stopped after drawing the hybrid height: empty donor window
            assert rec["result"] is None
            assert NM.add_reticulation_at(net, placements, k, rec["uniforms"][0], 0.5, 0.5) is None
            continue
        u2, u1, g = rec["uniforms"]
        out = NM.add_reticulation_at(net, placements, k, u2, u1, g)
        assert (out is None) == (rec["result"] is None)
        if out is None:
            continue
        new, loghr = out
        assert loghr == pytest.approx(rec["result"]["loghr"], abs=1e-9)
        got, want = _keyed(new, old), rec["result"]["network"]
        assert [n[0] for n in got["nodes"]] == [n[0] for n in want["nodes"]]
        assert [n[1] for n in got["nodes"]] == pytest.approx([n[1] for n in want["nodes"]], abs=1e-15)
        assert got["edges"] == want["edges"]
        assert new.n_reticulations() == net.n_reticulations() + 1


@pytest.mark.parametrize("case", [c for c in CASES if c["delete_moves"]], ids=[c["name"] for c in CASES if c["delete_moves"]])
def test_informed_delete_move_equals_the_reference(case):
    net = SpeciesNetwork.from_dict(case["network"])
    lab = net.node_label
    fn = _placements_fn(case)
    for rec in case["delete_moves"]:
        out = NM.delete_reticulation_at(net, lab.index(rec["retic"]), lab.index(rec["donor_parent"]), fn)
        assert out is not None
        new, loghr = out
        assert loghr == pytest.approx(rec["loghr"], abs=1e-9)
        got = _canonical(new)
        assert got["edges"] == rec["network"]["edges"]
        assert [n[0] for n in got["nodes"]] == [n[0] for n in rec["network"]["nodes"]]


def test_add_and_delete_are_exact_reverses():
    """log H(add: x -> x') == -log H(delete: x' -> x) when both use the same scoring rule -- the identity the
    reference's own tests check for its operators (tests/test_mcmc_seq_coupled.py)."""
    case = CASES[1]
    net = SpeciesNetwork.from_dict(case["network"])
    fn = _placements_fn(case)
    placements = fn(net)
    rng = np.random.default_rng(4)
    n_checked = 0
    for _ in range(8):
        k = int(rng.integers(len(placements)))
        out = NM.add_reticulation_at(net, placements, k, float(rng.random()), float(rng.random()), float(rng.uniform(0.1, 0.9)))
        if out is None:
            continue
        new, h_add = out
        v1, v2 = net.n_nodes, net.n_nodes + 1
        back = NM.delete_reticulation_at(new, v2, v1, fn)
        if back is None:
            continue
        old, h_del = back
        assert h_add == pytest.approx(-h_del, abs=1e-9)
        assert _canonical(old)["nodes"] == _canonical(net)["nodes"]
        n_checked += 1
    assert n_checked >= 4
    # the random-draw wrappers
    mv = NM.add_reticulation_move(net, placements, np.random.default_rng(1))
    assert mv is None or (mv[0].n_reticulations() == 1 and math.isfinite(mv[1]))
    assert NM.delete_reticulation_move(net, np.random.default_rng(1), fn) is None          # a tree has nothing to delete
    if mv is not None:
        dm = NM.delete_reticulation_move(mv[0], np.random.default_rng(2), fn)
        assert dm is None or dm[0].n_reticulations() == 0


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_scored_placements_with_the_real_prior(case):
    """The scan with the reference's own prior at its defaults (phynetpy_b200/priors.py as ``prior_fn``): the
    identifiability guards drop the same placements, every surviving score agrees."""
    from phynetpy_b200.priors import SeqPriors, log_prior_seq
    net = SpeciesNetwork.from_dict(case["network"])
    trees = [FlatTree(t["parent"], t["blen"], t["label"]) for t in case["trees"]]
    fn = _density_fn(case, NM.scaled_gene_trees(trees))
    pri = SeqPriors()
    got = NM.add_reticulation_placements(net, fn, len(HEIGHT_SCALE_GRID), prior_fn=lambda n: log_prior_seq(n, case["theta"], pri))
    mine = {(tuple(_labels(net, p["donor_edge"])), tuple(_labels(net, p["retic_edge"]))): p["s_rep"] for p in got}
    ref = {(tuple(p["donor"]), tuple(p["retic"])): p["s_rep"] for p in case["placements_with_prior"]}
    # Placements that split an in-edge of an EXISTING reticulation leave that node with one fresh in-edge of gamma 0.0
    # (the reference's _split_edge); its Beta prior is then evaluated on whichever in-edge an unordered set lists
    # first -- -inf or finite from run to run.  Those placements are compared only where both sides kept them.
    retic_in = {tuple(_labels(net, e)) for e in range(net.edge_src.shape[0])
                if (net.edge_dst == net.edge_dst[e]).sum() >= 2}
    ambiguous = {k for k in set(mine) | set(ref) if k[0] in retic_in or k[1] in retic_in}
    assert set(mine) - ambiguous == set(ref) - ambiguous
    assert len(ref) <= len(case["placements"])
    for k, v in ref.items():
        if k in mine:
            assert mine[k] == pytest.approx(v, rel=1e-11)
