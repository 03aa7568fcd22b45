"""The guided joint gene-tree re-proposal of the coupled moves with all scoring on the device
(phynetpy_b200/coupled.py) against the reference's ``_coupled_gene_tree_reproposal`` (golden vectors:
tests/golden/coupled_cases.json).  The member the reference chose is forced (its draw depends on the hash order
of a set of node objects); what is compared are the two selection log-densities it then computes."""
import json
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "coupled_cases.json")))["cases"]


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
def test_coupled_reproposal_matches_the_reference(case):
    from phynetpy_b200 import infer as P
    from phynetpy_b200.candidates import node_heights
    from phynetpy_b200.coupled import coupled_gene_tree_reproposal, _member
    model = P.GTR([0.1, 0.2, 0.3, 0.4], [1.0, 2.5, 0.7, 1.3, 3.1, 0.9])
    eng = P.SeqLikelihoodEngine(case["loci"], case["species_of"], model)
    trees = [P.FlatTree(t["parent"], t["blen"], t["label"]) for t in case["trees"]]
    heights = [node_heights(t) for t in trees]
    eng.log_likelihood(trees, None, 0.0)           # committed partials: candidates reuse them
    target = P.SpeciesNetwork.from_dict(case["target"])
    reverse = P.SpeciesNetwork.from_dict(case["reverse"])
    want = case["result"]
    assert want is not None
    picked = []

    def choose(i, table, probs):
        par, bl, ht = table
        for j in range(par.shape[0]):
            if _same(_describe(_member(trees[i], par[j], bl[j]), ht[j]), want["chosen"][i]):
                assert probs[j] > 0.0
                picked.append(j)
                return j
        raise AssertionError("the reference's choice for locus %d is not in the candidate set" % i)

    res = coupled_gene_tree_reproposal(eng, trees, heights, target, reverse, case["theta"], np.random.default_rng(0), choose)
    assert res is not None and len(picked) == len(trees)
    new_trees, new_heights, log_qf, log_qr = res
    for t, h, w in zip(new_trees, new_heights, want["chosen"]):
        assert _same(_describe(t, h), w)
        assert np.allclose(node_heights(t), h, rtol=0, atol=1e-12)
    assert log_qf == pytest.approx(want["log_qf"], abs=1e-9)
    assert log_qr == pytest.approx(want["log_qr"], abs=1e-9)


def test_coupled_reproposal_with_its_own_draws_is_reversible_in_distribution():
    """Own random draws: the move returns finite densities, never changes its inputs, and proposing back from
    the result (networks swapped) finds the original tree in its candidate set -- the symmetry the Hastings
    ratio relies on (src/_mcmc_seq.py:2009-2013)."""
    from phynetpy_b200 import infer as P
    from phynetpy_b200.candidates import node_heights, signature
    from phynetpy_b200.coupled import coupled_gene_tree_reproposal
    case = CASES[2]
    model = P.GTR([0.1, 0.2, 0.3, 0.4], [1.0, 2.5, 0.7, 1.3, 3.1, 0.9])
    eng = P.SeqLikelihoodEngine(case["loci"], case["species_of"], model)
    trees = [P.FlatTree(t["parent"], t["blen"], t["label"]) for t in case["trees"]]
    heights = [node_heights(t) for t in trees]
    eng.log_likelihood(trees, None, 0.0)
    target = P.SpeciesNetwork.from_dict(case["target"])
    reverse = P.SpeciesNetwork.from_dict(case["reverse"])
    sigs0 = [signature(t, h) for t, h in zip(trees, heights)]
    n_done = 0
    for seed in range(6):
        res = coupled_gene_tree_reproposal(eng, trees, heights, target, reverse, case["theta"], np.random.default_rng(seed))
        if res is None:
            continue
        n_done += 1
        nt, nh, qf, qr = res
        assert math.isfinite(qf) and math.isfinite(qr) and qf <= 1e-12 and qr <= 1e-12
        assert [signature(t, h) for t, h in zip(trees, heights)] == sigs0
        back = coupled_gene_tree_reproposal(eng, nt, nh, reverse, target, case["theta"], np.random.default_rng(seed + 100),
                                            choose=lambda i, table, probs: next(
                                                j for j in range(table[0].shape[0])
                                                if signature(P.FlatTree(table[0][j], table[1][j], trees[i].label), table[2][j]) == sigs0[i]))
        assert back is not None
        # forward density of the way back == reverse density of the way out (the very same candidate sets) ...
        assert back[2] == pytest.approx(qr, abs=1e-9)
        # ... and vice versa up to the re-derivation of branch lengths: the way back re-selects the original tree
        # as materialised from its heights, whose lengths differ from the stored ones in the 10th digit (the
        # reference's simulator prints 10 decimals), so the weights move by ~1e-8
        assert back[3] == pytest.approx(qf, abs=1e-6)
    assert n_done >= 3
