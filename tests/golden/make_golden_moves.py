"""Golden vectors for the gene-tree proposals, from the UNMODIFIED reference (``_slide_height_move``
``src/_mcmc_seq.py:956-990``, ``op_gene_tree_nni`` ``:1053-1113``, ``_sync_lengths`` ``:418-434``).
Build-container only (see _ref_loader.py).

    python tests/golden/make_golden_moves.py   ->  tests/golden/moves_cases.json

Per tree: (a) slide records -- the node the reference chose (as its leaf set), the uniform it drew and the
outcome (new height + log Hastings ratio, or null when it rejected); (b) the set of distinct trees
``op_gene_tree_nni`` produced over many seeds (each as sorted [leaf set, height] pairs of its internal
nodes plus every child->parent branch length), and whether it ever rejected.
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
from phynetpy.GraphUtils import _clone_net, _node_heights  # noqa: E402
from phynetpy.Network import Network  # noqa: E402
from phynetpy.models import BranchLengthUnit  # noqa: E402

from phynetpy_b200 import simulate as S  # noqa: E402  (input generator only)
from make_golden_candidates import newick  # noqa: E402
from make_golden_msnc import enewick, export_network, random_network  # noqa: E402

UNIT = BranchLengthUnit.SUBSTITUTIONS_PER_SITE


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


def describe(net, heights):
    nodes = [[clade(net, v), float(heights[v])] for v in net.V() if net.get_children(v)]
    nodes.sort(key=lambda x: (x[0], x[1]))
    lengths = []
    for e in net.E():
        lengths.append([clade(net, e.dest), clade(net, e.src), float(e.get_length())])
    lengths.sort()
    return {"internal": nodes, "lengths": lengths}


class _Engine:
    n_loci = 1

    def clone_caches(self):
        return ([], [])

    def invalidate_locus(self, i):
        pass

    def restore_caches(self, snap):
        pass


class _State:
    def __init__(self, gt):
        self._engine = _Engine()
        self.gene_trees = [gt]
        self.gt_heights = [_node_heights(gt)]


class _NetEngine(_Engine):
    def invalidate_network(self):
        pass


class _NetState:
    def __init__(self, net):
        self._engine = _NetEngine()
        self.species_net = net
        self.net_heights = _node_heights(net)


def network_cases(rng0):
    """Species-network parameter moves: _slide_height_move on networks (op_net_node_height) and op_change_gamma."""
    out = []
    for n_leaves, n_retic in ((4, 1), (6, 2), (8, 3), (5, 0)):
        nodes, edges = random_network(n_leaves, n_retic, rng0, height=0.08)
        nw = enewick(nodes, edges)
        net = Network.from_newick(nw, branch_length_unit=UNIT)
        slides = []
        for seed in range(50):
            work = _clone_net(net)
            heights = _node_heights(work)
            cands = M._internal_nodes(work) + [work.root()]
            replay = np.random.default_rng(seed)
            k = int(replay.integers(len(cands)))
            u = float(replay.random())
            loghr = M._slide_height_move(work, heights, np.random.default_rng(seed), 0.3)
            rec = {"seed": seed, "n_candidates": len(cands), "node": cands[k].label, "u": u}
            rec["result"] = None if loghr is None else {"height": float(heights[cands[k]]), "loghr": float(loghr)}
            slides.append(rec)
        gammas = []
        for seed in range(30 if n_retic else 2):
            st = _NetState(net)
            rets = [v for v in net.V() if v.is_reticulation()]
            res = M.op_change_gamma(st, np.random.default_rng(seed), window=0.6)
            if not rets:
                gammas.append({"seed": seed, "result": None})
                assert res is None
                continue
            replay = np.random.default_rng(seed)
            replay.integers(len(rets))
            u = float(replay.random())
            new = st.species_net
            changed = None
            for v in new.V():
                if not v.is_reticulation():
                    continue
                old_v = [x for x in net.V() if x.label == v.label][0]
                g_new = {e.src.label: float(e.get_gamma()) for e in new.in_edges(v)}
                g_old = {e.src.label: float(e.get_gamma()) for e in net.in_edges(old_v)}
                if g_new != g_old:
                    changed = {"retic": v.label, "old": g_old, "new": g_new}
            gammas.append({"seed": seed, "u": u, "window": 0.6, "n_reticulations": len(rets), "result": changed})
        out.append({"name": "net_n%d_r%d" % (n_leaves, n_retic), "newick": nw, "network": export_network(net),
                    "slides": slides, "gammas": gammas})
    return out


def main():
    rng0 = np.random.default_rng(77)
    cases = []
    for n in (3, 4, 5, 7, 10, 16):
        par, bl = S.random_trees(1, n, rng0, height=0.1)
        label = S.tip_labels(n) + [None] * (n - 1)
        nw = newick(par[0].tolist(), bl[0].tolist(), label)
        gt = Network.from_newick(nw, branch_length_unit=UNIT)
        slides = []
        for seed in range(60):
            work = _clone_net(gt)
            heights = _node_heights(work)
            cands = M._internal_nodes(work) + [work.root()]
            replay = np.random.default_rng(seed)
            k = int(replay.integers(len(cands)))
            u = float(replay.random())
            before = dict(heights)
            loghr = M._slide_height_move(work, heights, np.random.default_rng(seed), 0.3)
            rec = {"seed": seed, "n_candidates": len(cands), "node": clade(work, cands[k]), "u": u,
                   "is_root": cands[k] is work.root()}
            if loghr is None:
                rec["result"] = None
            else:
                changed = [v for v in work.V() if heights[v] != before[v]]
                assert len(changed) <= 1 and (not changed or changed[0] is cands[k])
                M._sync_lengths(work, heights)
                rec["result"] = {"height": float(heights[cands[k]]), "loghr": float(loghr)}
                if seed % 10 == 0:     # the whole re-synchronised tree for a sample of the records
                    rec["result"]["tree"] = describe(work, heights)
            slides.append(rec)
        outcomes, rejected, n_runs = {}, 0, 400
        for seed in range(n_runs):
            st = _State(gt)
            res = M.op_gene_tree_nni(st, np.random.default_rng(1000 + seed))
            if res is None:
                rejected += 1
                continue
            d = describe(st.gene_trees[0], st.gt_heights[0])
            outcomes[json.dumps(d, sort_keys=True)] = d
        cases.append({"name": "n%d" % n, "parent": par[0].tolist(), "blen": [None if b != b else b for b in bl[0].tolist()],
                      "label": label, "slides": slides, "nni_outcomes": list(outcomes.values()),
                      "nni_rejected": rejected, "nni_runs": n_runs})
    nets = network_cases(rng0)
    out = os.path.join(HERE, "moves_cases.json")
    with open(out, "w") as f:
        json.dump({"source": "PhyNetPy v0.6.0 _slide_height_move / op_gene_tree_nni / op_change_gamma", "cases": cases,
                   "network_cases": nets}, f)
    print("wrote %s: %d trees, %d slide records, %d distinct NNI outcomes" % (
        out, len(cases), sum(len(c["slides"]) for c in cases), sum(len(c["nni_outcomes"]) for c in cases)))


if __name__ == "__main__":
    main()
