"""Golden vectors for the guided joint gene-tree re-proposal of the coupled moves, from the UNMODIFIED reference
(``_coupled_gene_tree_reproposal`` ``src/_mcmc_seq.py:1995-2070``).  Build-container only (see _ref_loader.py).

    python tests/golden/make_golden_coupled.py   ->  tests/golden/coupled_cases.json

Each case: a species tree (the reverse network) and the same tree with one reticulation added (the target
network), per-locus alignments and current gene trees, theta, a seed -- and what the reference returned:
the chosen gene tree of every locus (as [leaf set, height] pairs of its internal nodes), log q_f and log q_r.
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

from make_golden_msnc import add_reticulation, enewick, export_gene_tree, export_network, random_network  # noqa: E402

UNIT = BranchLengthUnit.SUBSTITUTIONS_PER_SITE
PI, RATES = [0.1, 0.2, 0.3, 0.4], [1.0, 2.5, 0.7, 1.3, 3.1, 0.9]


class _State:
    pass


def clade(net, node):
    out, stack = [], [node]
    while stack:
        v = stack.pop()
        kids = net.get_children(v)
        if kids:
            stack.extend(kids)
        else:
            out.append(v.label)
    return sorted(out)


def describe(gt, heights):
    return sorted([clade(gt, v), float(heights[v])] for v in gt.V() if gt.get_children(v))


def coupled_move_cases(model):
    """op_add_reticulation_coupled / op_delete_reticulation_coupled with the sampler's prior patched to 0 and a
    recording generator: the placement (labels), the three uniforms, the chosen gene tree of every locus, the
    resulting network (new nodes named by height) and the log Hastings ratio."""
    import types
    from make_golden_netmoves import RecordingRng, canonical, keyed
    M.log_prior_seq = lambda *a, **k: 0.0
    out = []
    for n_species, n_retic, n_loci, L, seed in [(4, 0, 3, 60, 11), (5, 0, 2, 50, 12), (4, 1, 2, 60, 13), (5, 1, 3, 40, 14)]:
        rng = np.random.default_rng(1500 + seed)
        nodes, edges = random_network(n_species, n_retic, rng, height=0.06)
        nw = enewick(nodes, edges)
        net = Network.from_newick(nw, branch_length_unit=UNIT)
        alleles = {"s%d" % i: ["s%d" % i] for i in range(n_species)}
        species_of = {"s%d" % i: "s%d" % i for i in range(n_species)}
        theta = 0.03
        gts = [simulate_gene_tree(net, alleles, theta, rng) for _ in range(n_loci)]
        loci = [simulate_sequences(g, model, L, rng) for g in gts]

        def state():
            st = _State()
            st._engine = M._SeqLikelihoodEngine(loci, species_of, model, None)
            st.gene_trees = list(gts)
            st.gt_heights = [_node_heights(g) for g in gts]
            st.theta = theta
            st.model = model
            st.branch_thetas = None
            st.priors = types.SimpleNamespace(max_reticulations=10, max_level=None)
            st.species_net = net
            st.net_heights = _node_heights(net)
            st.num_reticulations = lambda: sum(1 for v in st.species_net.V() if v.is_reticulation())
            return st
        old_labels = {v.label for v in net.V()}
        recs = []
        for s_ in range(4):
            st = state()
            pl = M._add_reticulation_placements(st, net, _node_heights(net))
            rr = RecordingRng(2000 + 10 * seed + s_)
            res = M.op_add_reticulation_coupled(st, rr)
            if res is None:
                continue
            k = rr.log[0][1]
            recs.append({"kind": "add", "donor": list(pl[k]["donor_labels"]), "retic": list(pl[k]["retic_labels"]),
                         "uniforms": [x[1] for x in rr.log[1:4]], "loghr": float(res[0]),
                         "network": keyed(st.species_net, old_labels),
                         "chosen": [describe(g, h) for g, h in zip(st.gene_trees, st.gt_heights)]})
        for s_ in range(4 if n_retic else 0):
            st = state()
            res = M.op_delete_reticulation_coupled(st, RecordingRng(3000 + 10 * seed + s_))
            if res is None:
                continue
            gone = sorted(old_labels - {v.label for v in st.species_net.V()})
            retic = [x for x in gone if [v for v in net.V() if v.label == x][0].is_reticulation()][0]
            recs.append({"kind": "delete", "retic": retic, "donor_parent": [x for x in gone if x != retic][0],
                         "loghr": float(res[0]), "network": canonical(st.species_net),
                         "chosen": [describe(g, h) for g, h in zip(st.gene_trees, st.gt_heights)]})
        for s_ in range(8 if n_retic else 0):
            st = state()
            rr = RecordingRng(4000 + 10 * seed + s_)
            res = M.op_relocate_reticulation_coupled(st, rr)
            if res is None:
                continue
            new = st.species_net
            gone = sorted(old_labels - {v.label for v in new.V()})
            retic = [x for x in gone if [v for v in net.V() if v.label == x][0].is_reticulation()][0]
            fresh = [v for v in new.V() if v.label not in old_labels]
            nv2 = [v for v in fresh if v.is_reticulation()][0]
            nv1 = [v for v in fresh if v is not nv2][0]
            b3 = new.get_parents(nv1)[0]
            b4 = [c for c in new.get_children(nv1) if c is not nv2][0]
            b5 = [p for p in new.get_parents(nv2) if p is not nv1][0]
            b6 = new.get_children(nv2)[0]
            rand = [x[1] for x in rr.log if x[0] == "random"][:3]
            recs.append({"kind": "relocate", "retic": retic, "donor_parent": [x for x in gone if x != retic][0],
                         "donor": [b3.label, b4.label], "retic_edge": [b5.label, b6.label], "uniforms": rand,
                         "loghr": float(res[0]), "network": keyed(new, old_labels),
                         "chosen": [describe(g, h) for g, h in zip(st.gene_trees, st.gt_heights)]})
        out.append({"name": "n%d_r%d_m%d" % (n_species, n_retic, n_loci), "network": export_network(net), "theta": theta,
                    "loci": loci, "species_of": species_of, "trees": [export_gene_tree(g) for g in gts], "moves": recs})
        print(out[-1]["name"], [(r["kind"], round(r["loghr"], 3)) for r in recs])
    return out


def main():
    cases = []
    model = GTR(PI, RATES)
    for n_species, n_loci, L, seed in [(4, 3, 80, 1), (5, 4, 60, 2), (6, 3, 100, 3), (5, 5, 40, 4)]:
        rng = np.random.default_rng(500 + seed)
        nodes, edges = random_network(n_species, 0, rng, height=0.06)
        tree_nw = enewick(nodes, edges)
        # the same tree with one reticulation: re-run the generator's insertion on a copy
        rng_net = np.random.default_rng(900 + seed)
        nodes2 = [list(x) for x in nodes]
        edges2 = [list(x) for x in edges]

        add_reticulation(nodes2, edges2, rng_net, 0, gamma_range=(0.2, 0.8))
        net_nw = enewick(nodes2, edges2)
        sp_tree = Network.from_newick(tree_nw, branch_length_unit=UNIT)
        sp_net = Network.from_newick(net_nw, branch_length_unit=UNIT)
        alleles = {"s%d" % i: ["s%d" % i] for i in range(n_species)}
        species_of = {"s%d" % i: "s%d" % i for i in range(n_species)}
        theta = 0.03
        gts, loci = [], []
        for _ in range(n_loci):
            gt = simulate_gene_tree(sp_tree, alleles, theta, rng)
            gts.append(gt)
            loci.append(simulate_sequences(gt, model, L, rng))
        for direction, (target, reverse) in (("add", (sp_net, sp_tree)), ("delete", (sp_tree, sp_net))):
            st = _State()
            st._engine = M._SeqLikelihoodEngine(loci, species_of, model, None)
            st.gene_trees = list(gts)
            st.gt_heights = [_node_heights(g) for g in gts]
            st.theta = theta
            st.model = model
            st.branch_thetas = None
            res = M._coupled_gene_tree_reproposal(st, target, reverse, np.random.default_rng(7000 + seed))
            rec = {"name": "n%d_m%d_%s" % (n_species, n_loci, direction), "seed": 7000 + seed, "theta": theta,
                   "target": export_network(target), "reverse": export_network(reverse),
                   "loci": loci, "species_of": species_of, "trees": [export_gene_tree(g) for g in gts]}
            if res is None:
                rec["result"] = None
            else:
                new_gts, new_hs, log_qf, log_qr = res
                rec["result"] = {"chosen": [describe(g, h) for g, h in zip(new_gts, new_hs)],
                                 "log_qf": float(log_qf), "log_qr": float(log_qr)}
            cases.append(rec)
    moves = coupled_move_cases(model)
    out = os.path.join(HERE, "coupled_cases.json")
    with open(out, "w") as f:
        json.dump({"source": "PhyNetPy v0.6.0 _coupled_gene_tree_reproposal, op_add/delete_reticulation_coupled (prior at 0)",
                   "cases": cases, "move_cases": moves}, f)
    print("wrote %s: %d cases (%d declined)" % (out, len(cases), sum(c["result"] is None for c in cases)))
    for c in cases:
        print(c["name"], None if c["result"] is None else (round(c["result"]["log_qf"], 4), round(c["result"]["log_qr"], 4)))


if __name__ == "__main__":
    main()
