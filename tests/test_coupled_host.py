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
