"""Golden vectors for the reticulation surgery and the placement scan of the informed add / delete moves, from the
UNMODIFIED reference (``_geometric_placements``, ``_add_reticulation_placements``, ``_network_only_log_score``,
``_split_edge``, ``_suppress_degree2``; ``src/_mcmc_seq.py:1116-1362,1545-1655``).  Build-container only.

    python tests/golden/make_golden_netmoves.py   ->  tests/golden/netmoves_cases.json

The prior (``log_prior_seq``) belongs to the sampler and is out of scope: it is patched to 0 while the placements are
scored, so ``s_rep`` here is the MSNC part alone, sum_i logsumexp_s log P(g_i scaled by s | candidate network).
Nodes and edges are identified by labels (the reference's object order is not reproducible).
"""
import json
import math
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
from phynetpy._sim_seq import simulate_gene_tree  # noqa: E402
from phynetpy.GraphUtils import _clone_net, _node_heights  # noqa: E402
from phynetpy.Network import Network  # noqa: E402
from phynetpy.models import BranchLengthUnit  # noqa: E402

from make_golden_msnc import enewick, export_gene_tree, export_network, random_network  # noqa: E402

UNIT = BranchLengthUnit.SUBSTITUTIONS_PER_SITE


class _State:
    pass


class RecordingRng:
    """A numpy Generator that remembers what it returned, so the deterministic half of an operator can be replayed."""

    def __init__(self, seed):
        self.r = np.random.default_rng(seed)
        self.log = []

    def choice(self, n, p=None):
        k = int(self.r.choice(n, p=p))
        self.log.append(["choice", k])
        return k

    def random(self):
        u = float(self.r.random())
        self.log.append(["random", u])
        return u

    def integers(self, n):
        k = int(self.r.integers(n))
        self.log.append(["integers", k])
        return k


def keyed(net, old_labels):
    """Canonical form in which nodes the move created are named by their height (their labels are arbitrary)."""
    h = _node_heights(net)

    def key(v):
        return v.label if v.label in old_labels else "new@%r" % float(h[v])
    return {"nodes": sorted([key(v), float(h[v])] for v in net.V()),
            "edges": sorted([key(e.src), key(e.dest), None if e.get_gamma() is None else float(e.get_gamma())] for e in net.E())}


def make_state(net, gts, loci, species_of, model, theta):
    import types
    st = _State()
    st._engine = M._SeqLikelihoodEngine(loci, species_of, model, None)
    st.gene_trees = gts
    st.gt_heights = [_node_heights(g) for g in gts]
    st.theta = theta
    st.branch_thetas = None
    st.priors = types.SimpleNamespace(max_reticulations=10, max_level=None)
    st.species_net = net
    st.net_heights = _node_heights(net)
    st.num_reticulations = lambda: sum(1 for v in st.species_net.V() if v.is_reticulation())
    return st


def canonical(net):
    h = _node_heights(net)
    return {"nodes": sorted([v.label, float(h[v])] for v in net.V()),
            "edges": sorted([e.src.label, e.dest.label, None if e.get_gamma() is None else float(e.get_gamma())] for e in net.E())}


def delete_surgery(net, v2_label, v1_label):
    """The surgery lines of op_delete_reticulation (:1500-1567) without scoring."""
    old_heights = _node_heights(net)
    v2 = [v for v in net.V() if v.label == v2_label][0]
    parents = net.get_parents(v2)
    if len(parents) != 2:
        return None
    v1 = [p for p in parents if p.label == v1_label][0]
    v5 = parents[1] if parents[0] is v1 else parents[0]
    if v1.is_reticulation() or net.in_degree(v1) == 0:
        return None
    v3s = net.get_parents(v1)
    if len(v3s) != 1:
        return None
    v3 = v3s[0]
    other = [c for c in net.get_children(v1) if c is not v2]
    if len(other) != 1:
        return None
    v4 = other[0]
    v6s = net.get_children(v2)
    if len(v6s) != 1:
        return None
    v6 = v6s[0]
    if v3 in net.get_parents(v4) or v5 in net.get_parents(v6) or (v3 is v5 and v4 is v6):
        return None
    new_net = _clone_net(net)
    new_h = _node_heights(new_net)
    cv2 = [v for v in new_net.V() if v.label == v2_label][0]
    cv1 = [p for p in new_net.get_parents(cv2) if p.label == v1_label][0]
    new_net.remove_edge(M._edge_between(new_net, cv1, cv2))
    cv2.set_is_reticulation(False)
    M._suppress_degree2(new_net, new_h, cv1)
    M._suppress_degree2(new_net, new_h, cv2)
    if len(list(new_net.E())) < 2:
        return None
    M._sync_lengths(new_net, new_h)
    return canonical(new_net)


def main():
    orig_prior = M.log_prior_seq
    M.log_prior_seq = lambda *a, **k: 0.0          # scores without the prior (the MSNC part alone)
    model = GTR([0.25] * 4, [1.0] * 6)
    cases = []
    for n_species, n_retic, n_loci, seed in [(3, 0, 2, 1), (4, 0, 3, 2), (5, 0, 2, 3), (4, 1, 2, 4), (5, 1, 2, 5), (5, 2, 2, 6)]:
        rng = np.random.default_rng(300 + seed)
        nodes, edges = random_network(n_species, n_retic, rng, height=0.06)
        nw = enewick(nodes, edges)
        net = Network.from_newick(nw, branch_length_unit=UNIT)
        heights = _node_heights(net)
        alleles = {"s%d" % i: ["s%d" % i] for i in range(n_species)}
        species_of = {"s%d" % i: "s%d" % i for i in range(n_species)}
        theta = 0.03
        gts = [simulate_gene_tree(net, alleles, theta, rng) for _ in range(n_loci)]
        loci = [{("s%d" % i): "ACGT" for i in range(n_species)} for _ in range(n_loci)]
        st = _State()
        st._engine = M._SeqLikelihoodEngine(loci, species_of, model, None)
        st.gene_trees = gts
        st.theta = theta
        st.branch_thetas = None
        st.priors = None
        geo = [[list(p["donor_labels"]), list(p["retic_labels"])] for p in M._geometric_placements(net, heights)]
        placed = [{"donor": list(p["donor_labels"]), "retic": list(p["retic_labels"]), "l2": float(p["l2"]), "s_rep": float(p["s_rep"])}
                  for p in M._add_reticulation_placements(st, net, heights)]
        base_score = float(M._network_only_log_score(st, net))
        # ... and once with the reference's real prior at its default hyper-parameters (its identifiability guards
        # drop the placements that would create cycles of fewer than 4 nodes)
        M.log_prior_seq = orig_prior
        st.priors = M.MCMCSeqPriors()
        placed_prior = [{"donor": list(p["donor_labels"]), "retic": list(p["retic_labels"]), "s_rep": float(p["s_rep"])}
                        for p in M._add_reticulation_placements(st, net, heights)]
        M.log_prior_seq = lambda *a, **k: 0.0
        st.priors = None
        deletes = []
        for v in net.V():
            if v.is_reticulation():
                for p in net.get_parents(v):
                    deletes.append({"retic": v.label, "donor_parent": p.label, "result": delete_surgery(net, v.label, p.label)})
        # the informed moves themselves, with a recording generator (network part only: the prior is 0)
        old_labels = {v.label for v in net.V()}
        adds = []
        for s_ in range(6):
            st2 = make_state(net, gts, loci, species_of, model, theta)
            pl = M._add_reticulation_placements(st2, net, _node_heights(net))
            rr = RecordingRng(40 + s_)
            res = M.op_add_reticulation(st2, rr)
            k = rr.log[0][1]
            rec = {"donor": list(pl[k]["donor_labels"]), "retic": list(pl[k]["retic_labels"]),
                   "uniforms": [x[1] for x in rr.log[1:]]}
            rec["result"] = None if res is None else {"loghr": float(res[0]), "network": keyed(st2.species_net, old_labels)}
            adds.append(rec)
        dels = []
        for s_ in range(6 if n_retic else 0):
            st2 = make_state(net, gts, loci, species_of, model, theta)
            res = M.op_delete_reticulation(st2, RecordingRng(60 + s_))
            if res is None:
                continue            # which pair was drawn cannot be read back from a declined move
            gone = sorted(old_labels - {v.label for v in st2.species_net.V()})
            retic = [x for x in gone if [v for v in net.V() if v.label == x][0].is_reticulation()][0]
            donor_parent = [x for x in gone if x != retic][0]
            dels.append({"retic": retic, "donor_parent": donor_parent, "loghr": float(res[0]),
                         "network": canonical(st2.species_net)})
        cases.append({"name": "n%d_r%d" % (n_species, n_retic), "newick": nw, "network": export_network(net), "theta": theta,
                      "species_of": species_of, "trees": [export_gene_tree(g) for g in gts], "base_score": base_score,
                      "geometric": geo, "placements": placed, "placements_with_prior": placed_prior, "deletes": deletes,
                      "add_moves": adds, "delete_moves": dels})
        print(cases[-1]["name"], "edges", len(list(net.E())), "geometric", len(geo), "scored", len(placed), "deletes",
              [d["result"] is not None for d in deletes], "base", round(base_score, 4) if math.isfinite(base_score) else base_score)
    out = os.path.join(HERE, "netmoves_cases.json")
    with open(out, "w") as f:
        json.dump({"source": "PhyNetPy v0.6.0 informed add/delete reticulation helpers; prior patched to 0", "cases": cases}, f)
    print("wrote", out)


if __name__ == "__main__":
    main()
