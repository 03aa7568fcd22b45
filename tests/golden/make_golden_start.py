"""Golden vectors for the starting state, from the UNMODIFIED reference (``build_upgma_gene_tree``, ``_jc_distance``,
``_enforce_embedding_consistency``; ``src/_mcmc_seq.py:473-641``).  Build-container only.

    python tests/golden/make_golden_start.py   ->  tests/golden/start_cases.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _ref_loader import load_reference  # noqa: E402

load_reference()
from phynetpy import _mcmc_seq as M  # noqa: E402
from phynetpy._seq_likelihood import GTR  # noqa: E402
from phynetpy._sim_seq import simulate_gene_tree, simulate_sequences  # noqa: E402
from phynetpy.GraphUtils import _node_heights  # noqa: E402
from phynetpy.Network import Network  # noqa: E402
from phynetpy.models import BranchLengthUnit  # noqa: E402

from make_golden_msnc import enewick, export_gene_tree, export_network, random_network  # noqa: E402

UNIT = BranchLengthUnit.SUBSTITUTIONS_PER_SITE


def describe(gt):
    h = _node_heights(gt)

    def clade(v):
        out, stack = [], [v]
        while stack:
            x = stack.pop()
            kids = gt.get_children(x)
            if kids:
                stack.extend(kids)
            else:
                out.append(x.label)
        return sorted(out)
    nodes = sorted([clade(v), float(h[v])] for v in gt.V() if gt.get_children(v))
    lengths = sorted([clade(e.dest), float(e.get_length())] for e in gt.E())
    return {"internal": nodes, "lengths": lengths}


def main():
    rng = np.random.default_rng(2025)
    model = GTR([0.1, 0.2, 0.3, 0.4], [1.0, 2.5, 0.7, 1.3, 3.1, 0.9])
    cases = []
    for n_species, per_sp, L, theta, n_retic in [(3, 1, 80, 0.02, 0), (5, 1, 120, 0.03, 0), (4, 2, 100, 0.05, 0), (6, 1, 60, 0.02, 1),
                                                  (8, 1, 40, 0.04, 0), (5, 1, 30, 0.3, 1)]:
        nodes, edges = random_network(n_species, n_retic, rng, height=0.08)
        nw = enewick(nodes, edges)
        net = Network.from_newick(nw, branch_length_unit=UNIT)
        alleles = {"s%d" % i: (["s%d" % i] if per_sp == 1 else ["s%d_%d" % (i, k) for k in range(per_sp)]) for i in range(n_species)}
        species_of = {a: sp for sp, al in alleles.items() for a in al}
        loci = []
        for _ in range(3):
            gt = simulate_gene_tree(net, alleles, theta, rng)
            aln = simulate_sequences(gt, model, L, rng)
            loci.append(aln)
        # blemishes: lower case, gaps, an ambiguity code, one all-gap pair
        aln0 = dict(loci[0])
        k0, k1 = list(aln0)[:2]
        aln0[k0] = aln0[k0][:5].lower() + "-N" + aln0[k0][7:]
        aln0[k1] = "R" + aln0[k1][1:]
        loci[0] = aln0
        upgma = [M.build_upgma_gene_tree(a) for a in loci]
        dists = []
        for a in loci:
            labs = list(a)
            dists.append([[M._jc_distance(a[x], a[y]) for y in labs] for x in labs])
        net_h = _node_heights(net)
        before = {v.label: float(net_h[v]) for v in net.V()}
        scale = M._enforce_embedding_consistency(net, net_h, upgma, [_node_heights(g) for g in upgma], species_of)
        after = {v.label: float(net_h[v]) for v in net.V()}
        cases.append({"name": "n%d_a%d_r%d" % (n_species, per_sp, n_retic), "newick": nw, "network": export_network(Network.from_newick(nw, branch_length_unit=UNIT)),
                      "species_of": species_of, "loci": loci, "distances": dists, "upgma": [describe(g) for g in upgma],
                      "upgma_flat": [export_gene_tree(g) for g in upgma], "scale": float(scale), "heights_before": before,
                      "heights_after": after})
        print(cases[-1]["name"], "scale", round(scale, 4))
    cases.append({"name": "degenerate", "pairs": [["ACGT", "ACGT", M._jc_distance("ACGT", "ACGT")], ["----", "ACGT", M._jc_distance("----", "ACGT")],
                                                  ["AAAA", "CCCC", M._jc_distance("AAAA", "CCCC")], ["ACGTAC", "ACG", M._jc_distance("ACGTAC", "ACG")],
                                                  ["acgt", "ACGA", M._jc_distance("acgt", "ACGA")]]})
    out = os.path.join(HERE, "start_cases.json")
    with open(out, "w") as f:
        json.dump({"source": "PhyNetPy v0.6.0 build_upgma_gene_tree / _enforce_embedding_consistency", "cases": cases}, f)
    print("wrote", out)


if __name__ == "__main__":
    main()
