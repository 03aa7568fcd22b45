"""Host logic of the guided joint gene-tree re-proposal (phynetpy_b200/coupled.py: candidate tables, selection
densities, materialisation, reverse matching) on the CPU: the device scoring is replaced by the two oracles behind
the same two engine methods, and the result is compared with the reference's ``_coupled_gene_tree_reproposal``
(tests/golden/coupled_cases.json).  The GPU twin of this test is tests/test_gpu_coupled.py."""
import json
import os

import numpy as np
import pytest

from oracle import felsenstein_oracle as O, msnc_oracle as MO
from phynetpy_b200 import coupled as CP
from phynetpy_b200.candidates import candidate_arrays, candidate_table, node_heights, signature
from phynetpy_b200.tree import FlatTree

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "coupled_cases.json")))["cases"]


class OracleEngine:
    """The two scoring calls of SeqLikelihoodEngine, answered by the oracles."""

    def __init__(self, case):
        self.case = case
        self.n_loci = len(case["loci"])
        self.loc = [O.OracleLocus(d) for d in case["loci"]]
        self.model = O.gtr([0.1, 0.2, 0.3, 0.4], [1.0, 2.5, 0.7, 1.3, 3.1, 0.9])
        self.calls = 0

    def score_batch(self, batch, loci):
        self.calls += 1
        out = np.empty(batch.n_members)
        for m in range(batch.n_members):
            j = int(batch.tree_of_member[m])
            par, bl, _ = batch.member(m)
            out[m] = O.log_likelihood(self.loc[loci[j]], O.OracleTree(par.tolist(), [None if b != b else b for b in bl.tolist()],
                                                                      batch.trees[j].label), self.model)
        return out

    def msnc_batch(self, batch, loci, net, theta):
        self.calls += 1
        net = net if isinstance(net, dict) else net.to_dict()
        out = np.empty(batch.n_members)
        for m in range(batch.n_members):
            j = int(batch.tree_of_member[m])
            par, bl, _ = batch.member(m)
            out[m] = MO.msnc_log_density(net, par.tolist(), bl.tolist(), batch.trees[j].label, self.case["species_of"], theta)
        return out


def _describe(ft, h):
    kids = [[] for _ in range(ft.n_nodes)]
    for i, p in enumerate(ft.parent.tolist()):
        if p >= 0:
            kids[p].append(i)
    memo = {}

    def clade(v):
        if v not in memo:
            memo[v] = sorted(sum((clade(c) for c in kids[v]), [])) if kids[v] else [ft.label[v]]
        return memo[v]
    return sorted([clade(v), float(h[v])] for v in range(ft.n_nodes) if kids[v])


def _same(a, b):
    return len(a) == len(b) and all(x[0] == y[0] and abs(x[1] - y[1]) <= 1e-12 * max(1.0, abs(y[1])) for x, y in zip(a, b))


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_selection_densities_equal_the_reference(case):
    eng = OracleEngine(case)
    trees = [FlatTree(t["parent"], t["blen"], t["label"]) for t in case["trees"]]
    heights = [node_heights(t) for t in trees]
    want = case["result"]

    def choose(i, table, probs):
        par, bl, ht = table
        hits = [j for j in range(par.shape[0]) if _same(_describe(CP._member(trees[i], par[j], bl[j]), ht[j]), want["chosen"][i])]
        assert len(hits) == 1 and probs[hits[0]] > 0.0
        return hits[0]

    res = CP.coupled_gene_tree_reproposal(eng, trees, heights, case["target"], case["reverse"], case["theta"],
                                          np.random.default_rng(0), choose)
    assert res is not None
    new_trees, new_heights, log_qf, log_qr = res
    assert eng.calls == 4                                            # two scoring calls forward, two reverse
    for t, h, w in zip(new_trees, new_heights, want["chosen"]):
        assert _same(_describe(t, h), w)
    assert log_qf == pytest.approx(want["log_qf"], abs=1e-11)
    assert log_qr == pytest.approx(want["log_qr"], abs=1e-11)


def test_candidate_table_extends_candidate_arrays():
    for case in CASES[:3]:
        for t in case["trees"]:
            ft = FlatTree(t["parent"], t["blen"], t["label"])
            a, b = candidate_arrays(ft), candidate_table(ft)
            assert np.array_equal(a[0], b[0]) and np.array_equal(np.nan_to_num(a[1]), np.nan_to_num(b[1]))
            sigs = set()
            for k in range(b[0].shape[0]):
                c = FlatTree(b[0][k], b[1][k], ft.label)
                assert np.allclose(node_heights(c), b[2][k], rtol=0, atol=1e-12)
                sigs.add(signature(c, b[2][k]))
            assert len(sigs) == b[0].shape[0]                        # de-duplicated


def test_declines_like_the_reference_when_nothing_has_weight():
    case = CASES[0]
    eng = OracleEngine(case)
    eng.msnc_batch = lambda batch, loci, net, theta: np.full(batch.n_members, -np.inf)
    trees = [FlatTree(t["parent"], t["blen"], t["label"]) for t in case["trees"]]
    heights = [node_heights(t) for t in trees]
    assert CP.coupled_gene_tree_reproposal(eng, trees, heights, case["target"], case["reverse"], case["theta"],
                                           np.random.default_rng(0)) is None
    with pytest.raises(ValueError, match="one gene tree"):
        CP.coupled_gene_tree_reproposal(eng, trees[:1], heights, case["target"], case["reverse"], case["theta"],
                                        np.random.default_rng(0))


# ---------------------------------------------------------------------------------------------------
# the coupled reticulation moves: network surgery + joint gene-tree re-proposal
# ---------------------------------------------------------------------------------------------------
MOVE_CASES = json.load(open(os.path.join(HERE, "golden", "coupled_cases.json")))["move_cases"]


def _setup_move_case(case):
    from phynetpy_b200 import netmoves as NM
    from phynetpy_b200._msnc_density import SpeciesNetwork
    from phynetpy_b200.candidates import HEIGHT_SCALE_GRID
    eng = OracleEngine(case)
    net = SpeciesNetwork.from_dict(case["network"])
    trees = [FlatTree(t["parent"], t["blen"], t["label"]) for t in case["trees"]]
    heights = [node_heights(t) for t in trees]

    def placements_for(network, gene_trees):
        scaled = NM.scaled_gene_trees(gene_trees)

        def dens(cand):
            d = cand.to_dict()
            return np.array([MO.msnc_log_density(d, t.parent.tolist(), t.blen.tolist(), t.label, case["species_of"], case["theta"])
                             for t in scaled])
        return NM.add_reticulation_placements(network, dens, len(HEIGHT_SCALE_GRID))
    return eng, net, trees, heights, placements_for


def _chooser(trees, want):
    def choose(i, table, probs):
        par, bl, ht = table
        hits = [j for j in range(par.shape[0]) if _same(_describe(CP._member(trees[i], par[j], bl[j]), ht[j]), want[i])]
        assert len(hits) == 1
        return hits[0]
    return choose


def _keyed(net, old_labels):
    def key(v):
        lab = net.node_label[v]
        return lab if lab in old_labels else "new@%r" % float(net.height[v])
    return sorted([key(int(s)), key(int(d)), None if g != g else float(g)]
                  for s, d, g in zip(net.edge_src, net.edge_dst, net.edge_gamma.tolist()))


@pytest.mark.parametrize("case", MOVE_CASES, ids=[c["name"] for c in MOVE_CASES])
def test_coupled_reticulation_moves_equal_the_reference(case):
    """op_add_reticulation_coupled / op_delete_reticulation_coupled (prior at 0): with the reference's placement,
    uniforms and chosen gene trees forced, the same network and the same log Hastings ratio."""
    eng, net, trees, heights, placements_for = _setup_move_case(case)
    lab = net.node_label
    old = set(lab)
    placements = placements_for(net, trees)
    index = {((lab[int(net.edge_src[p["donor_edge"]])], lab[int(net.edge_dst[p["donor_edge"]])]),
              (lab[int(net.edge_src[p["retic_edge"]])], lab[int(net.edge_dst[p["retic_edge"]])])): k
             for k, p in enumerate(placements)}
    n_add = n_del = 0
    for rec in case["moves"]:
        choose = _chooser(trees, rec["chosen"])
        if rec["kind"] == "add":
            k = index[(tuple(rec["donor"]), tuple(rec["retic"]))]
            out = CP.coupled_add_reticulation(eng, net, trees, heights, case["theta"], np.random.default_rng(0), placements_for,
                                              placement=(k,) + tuple(rec["uniforms"]), choose=choose)
            assert out is not None
            new_net, new_trees, new_h, loghr = out
            assert _keyed(new_net, old) == rec["network"]["edges"]
            n_add += 1
        elif rec["kind"] == "delete":
            out = CP.coupled_delete_reticulation(eng, net, trees, heights, case["theta"], np.random.default_rng(0), placements_for,
                                                 pick=(lab.index(rec["retic"]), lab.index(rec["donor_parent"])), choose=choose)
            assert out is not None
            new_net, new_trees, new_h, loghr = out
            assert _keyed(new_net, old) == rec["network"]["edges"]
            n_del += 1
        else:       # relocate: detach A, re-attach at the placement B the reference drew on the base network
            from phynetpy_b200 import netmoves as NM
            pick = (lab.index(rec["retic"]), lab.index(rec["donor_parent"]))
            n0 = NM.delete_reticulation_geometry(net, *pick)["network"]
            l0 = n0.node_label
            edge = {(l0[int(s)], l0[int(d)]): e for e, (s, d) in enumerate(zip(n0.edge_src, n0.edge_dst))}
            d_e, r_e = edge[tuple(rec["donor"])], edge[tuple(rec["retic_edge"])]
            assert (d_e, r_e) in NM.geometric_placements(n0)
            out = CP.coupled_relocate_reticulation(eng, net, trees, heights, case["theta"], np.random.default_rng(0), pick=pick,
                                                   placement=(d_e, r_e) + tuple(rec["uniforms"]), choose=choose)
            assert out is not None
            new_net, new_trees, new_h, loghr = out
            assert _keyed(new_net, old) == rec["network"]["edges"]
            assert new_net.n_reticulations() == net.n_reticulations()
        assert loghr == pytest.approx(rec["loghr"], abs=1e-8)
        for t, h, w in zip(new_trees, new_h, rec["chosen"]):
            assert _same(_describe(t, h), w)
    assert n_add >= 4


def test_relocation_is_its_own_reverse():
    """log H(relocate A -> B) == -log H(relocate B -> A) when the way back re-attaches at A's place with A's heights
    and gamma and every locus re-selects its original tree (``_relocation_loghr`` is antisymmetric, ``:2578-2601``)."""
    from phynetpy_b200 import netmoves as NM
    from phynetpy_b200.candidates import signature
    case = next(c for c in MOVE_CASES if any(m["kind"] == "relocate" for m in c["moves"]))
    eng, net, trees, heights, placements_for = _setup_move_case(case)
    trees = [CP._member(t, t.parent, CP._lengths_from_heights(t.parent, h)) for t, h in zip(trees, heights)]
    sigs = [signature(t, h) for t, h in zip(trees, heights)]
    lab = net.node_label
    rng = np.random.default_rng(3)
    done = 0
    for attempt in range(10):
        pick = NM.draw_reticulation_to_delete(net, rng)
        fwd = CP.coupled_relocate_reticulation(eng, net, trees, heights, case["theta"], rng, pick=pick)
        if fwd is None:
            continue
        new_net, new_trees, new_h, h_fwd = fwd
        # the way back: detach B (the two nodes the forward move appended), re-attach on A's merged edges
        geom_a = NM.delete_reticulation_geometry(net, *pick)
        n0 = geom_a["network"]
        b_pick = (new_net.n_nodes - 1, new_net.n_nodes - 2)
        n0_back = NM.delete_reticulation_geometry(new_net, *b_pick)["network"]
        l0 = n0_back.node_label
        edge = {(l0[int(s)], l0[int(d)]): e for e, (s, d) in enumerate(zip(n0_back.edge_src, n0_back.edge_dst))}
        d_e, r_e = edge[geom_a["donor_labels"]], edge[geom_a["retic_labels"]]
        v2, v1 = pick
        h = net.height
        (b3, b4), (b5, b6) = [(lab.index(a), lab.index(b)) for a, b in (geom_a["donor_labels"], geom_a["retic_labels"])]
        u2 = (h[v2] - h[b6]) / (h[b5] - h[b6])
        lo = max(h[b4], h[v2])
        u1 = (h[v1] - lo) / (h[b3] - lo)
        e_hyb = int(np.nonzero((net.edge_src == v1) & (net.edge_dst == v2))[0][0])
        back = CP.coupled_relocate_reticulation(
            eng, new_net, new_trees, new_h, case["theta"], rng, pick=b_pick,
            placement=(d_e, r_e, float(u2), float(u1), float(net.edge_gamma[e_hyb])),
            choose=lambda i, table, probs: next(j for j in range(table[0].shape[0])
                                                if signature(CP._member(new_trees[i], table[0][j], table[1][j]), table[2][j]) == sigs[i]))
        if back is None:
            continue
        assert h_fwd + back[3] == pytest.approx(0.0, abs=1e-8)
        assert np.sort(back[0].height) == pytest.approx(np.sort(net.height), abs=1e-12)
        done += 1
    assert done >= 2


def test_coupled_add_and_delete_are_exact_reverses():
    """log H(add: x -> x') + log H(delete: x' -> x) == 0 when the delete picks the reticulation the add created and
    every locus re-selects its original gene tree -- the identity of tests/test_mcmc_seq_coupled.py in the reference,
    up to the re-derivation of branch lengths of the materialised trees (the simulator prints 10 decimals)."""
    from phynetpy_b200.candidates import signature
    case = MOVE_CASES[0]
    eng, net, trees, heights, placements_for = _setup_move_case(case)
    # start from trees whose lengths agree with their heights, so that materialising changes nothing
    trees = [CP._member(t, t.parent, CP._lengths_from_heights(t.parent, h)) for t, h in zip(trees, heights)]
    rng = np.random.default_rng(8)
    done = 0
    for attempt in range(6):
        fwd = CP.coupled_add_reticulation(eng, net, trees, heights, case["theta"], rng, placements_for)
        if fwd is None:
            continue
        new_net, new_trees, new_h, h_add = fwd
        sigs = [signature(t, h) for t, h in zip(trees, heights)]

        def back_choose(i, table, probs):
            par, bl, ht = table
            return next(j for j in range(par.shape[0]) if signature(CP._member(new_trees[i], par[j], bl[j]), ht[j]) == sigs[i])
        back = CP.coupled_delete_reticulation(eng, new_net, new_trees, new_h, case["theta"], rng, placements_for,
                                              pick=(net.n_nodes + 1, net.n_nodes), choose=back_choose)
        if back is None:
            continue
        old_net, old_trees, old_h, h_del = back
        assert h_add + h_del == pytest.approx(0.0, abs=1e-8)
        assert old_net.n_reticulations() == net.n_reticulations()
        done += 1
    assert done >= 2


def test_level_cap_rejects_before_the_reproposal():
    case = next(c for c in MOVE_CASES if sum(c["network"]["is_retic"]) >= 1)
    eng, net, trees, heights, placements_for = _setup_move_case(case)
    calls0 = eng.calls
    rng = np.random.default_rng(2)
    # every added reticulation lands in or next to the existing blob or makes its own: with a cap of 0 nothing passes
    out = CP.coupled_add_reticulation(eng, net, trees, heights, case["theta"], rng, placements_for, max_level=0)
    assert out is None and eng.calls == calls0            # declined without a single scoring call
    out = CP.coupled_relocate_reticulation(eng, net, trees, heights, case["theta"], rng, max_level=0)
    assert out is None and eng.calls == calls0
    out = CP.coupled_add_reticulation(eng, net, trees, heights, case["theta"], rng, placements_for, max_level=5)
    assert out is None or out[0].n_reticulations() == net.n_reticulations() + 1
