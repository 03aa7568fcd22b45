"""Starting state on flat objects (phynetpy_b200/start.py) against the reference's ``build_upgma_gene_tree``,
``_jc_distance`` and ``_enforce_embedding_consistency`` (tests/golden/start_cases.json)."""
import json
import os

import numpy as np
import pytest

from phynetpy_b200 import start as ST
from phynetpy_b200._msnc_density import SpeciesNetwork
from phynetpy_b200.candidates import node_heights
from phynetpy_b200.tree import FlatTree

HERE = os.path.dirname(os.path.abspath(__file__))
ALL = json.load(open(os.path.join(HERE, "golden", "start_cases.json")))["cases"]
CASES = [c for c in ALL if "loci" in c]


def _describe(ft):
    h = node_heights(ft)
    kids = [[] for _ in range(ft.n_nodes)]
    for i, p in enumerate(ft.parent.tolist()):
        if p >= 0:
            kids[p].append(i)
    memo = {}

    def clade(v):
        if v not in memo:
            memo[v] = sorted(sum((clade(c) for c in kids[v]), [])) if kids[v] else [ft.label[v]]
        return memo[v]
    return {"internal": sorted([clade(v), float(h[v])] for v in range(ft.n_nodes) if kids[v]),
            "lengths": sorted([clade(i), float(ft.blen[i])] for i in range(ft.n_nodes) if ft.parent[i] >= 0)}


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_jc_distances_and_upgma_equal_the_reference(case):
    for aln, want_d, want_t in zip(case["loci"], case["distances"], case["upgma"]):
        labels, d = ST.jc_distance_matrix(aln)
        assert labels == list(aln)
        assert np.array_equal(d, np.array(want_d))                    # same integer counts, same libm call
        ft = ST.build_upgma_gene_tree(aln)
        got = _describe(ft)
        assert [x[0] for x in got["internal"]] == [x[0] for x in want_t["internal"]]
        assert [x[1] for x in got["internal"]] == pytest.approx([x[1] for x in want_t["internal"]], abs=1e-15)
        assert got["lengths"] == want_t["lengths"]                   # ten-decimal lengths, to the bit
        assert sorted(l for l in ft.label if l and not l.startswith("I")) == sorted(aln)


def test_degenerate_pairs():
    deg = next(c for c in ALL if c["name"] == "degenerate")
    for a, b, want in deg["pairs"]:
        _, d = ST.jc_distance_matrix({"x": a, "y": b})
        assert d[0, 1] == want and d[1, 0] == want
    assert ST.build_upgma_gene_tree({"only": "ACGT"}).n_nodes == 1
    with pytest.raises(ValueError):
        ST.build_upgma_gene_tree({})


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_embedding_consistency_equals_the_reference(case):
    net = SpeciesNetwork.from_dict(case["network"])
    trees = [FlatTree(t["parent"], t["blen"], t["label"]) for t in case["upgma_flat"]]
    heights = [node_heights(t) for t in trees]
    new, scale = ST.enforce_embedding_consistency(net, trees, heights, case["species_of"])
    assert scale == pytest.approx(case["scale"], rel=1e-13, abs=1e-300)
    lab = net.node_label
    for v in range(net.n_nodes):
        assert float(net.height[v]) == pytest.approx(case["heights_before"][lab[v]], abs=1e-15)
        assert float(new.height[v]) == pytest.approx(case["heights_after"][lab[v]], rel=1e-13, abs=1e-300)
    if scale < 1.0 and scale > 0.0:
        # after scaling every gene tree embeds: the oracle density is finite
        from oracle import msnc_oracle as MO
        for t in trees:
            assert np.isfinite(MO.msnc_log_density(new.to_dict(), t.parent.tolist(), t.blen.tolist(), t.label, case["species_of"], 0.03))
