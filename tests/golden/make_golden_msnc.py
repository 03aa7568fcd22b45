"""Golden vectors for the timed MSNC density log P(g | Psi), from the UNMODIFIED reference
(``gene_tree_msnc_log_density`` ``src/_msnc_density.py:496-526`` -> ``_msnc_log_density_timed``
``:213-389`` / ``network_dp_timed_cy`` ``src/cython/gt_msc_cy.pyx:458-676``).  Build-container only
(see _ref_loader.py).

    python tests/golden/make_golden_msnc.py   ->  tests/golden/msnc_cases.json

(Node and edge numbering follow the reference's ``V()`` / ``E()``, sets of objects whose iteration order changes
from run to run: a regenerated file holds the same cases in a different numbering.)

Each case holds a species network as flat arrays read back from the reference's own
``_NetworkIndex`` (edge order, reticulation flags, gammas, node heights), a batch of gene trees
as flat parent / length / label arrays (reference ``V()`` order), the allele -> species map,
theta and the reference's density for every gene tree.  Gene trees come from the reference's own
simulator (``simulate_gene_tree`` ``src/_sim_seq.py:278``), plus height-rescaled copies -- most of
which no longer embed, pinning the ``-inf`` convention (``msnc_log_density_prebuilt`` ``:488-490``).
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
from phynetpy._msnc_density import (  # noqa: E402
    build_network_msnc_index,
    gene_tree_msnc_log_density,
)
from phynetpy._sim_seq import simulate_gene_tree  # noqa: E402
from phynetpy.Network import Network  # noqa: E402
from phynetpy.models import BranchLengthUnit  # noqa: E402

UNIT = BranchLengthUnit.SUBSTITUTIONS_PER_SITE


# ------------------------------------------------------------------ random ultrametric networks
def random_network(n_leaves, n_retic, rng, height=0.05):
    """(nodes, edges): nodes = [label, height, is_retic]; edges = [src, dst, gamma|None]."""
    nodes = [["s%d" % i, 0.0, False] for i in range(n_leaves)]
    edges = []
    live = list(range(n_leaves))
    t = 0.0
    while len(live) > 1:
        t += rng.exponential(height / n_leaves)
        i, j = sorted(rng.choice(len(live), size=2, replace=False).tolist(), reverse=True)
        a, b = live.pop(i), live.pop(j)
        nodes.append(["i%d" % len(nodes), t, False])
        v = len(nodes) - 1
        edges += [[v, a, None], [v, b, None]]
        live.append(v)
    for k in range(n_retic):
        add_reticulation(nodes, edges, rng, k)
    return nodes, edges


def add_reticulation(nodes, edges, rng, k=0, gamma_range=(0.1, 0.9)):
    """Insert hybrid node #Hk on a random recipient edge and its second parent xk on a donor edge at or above it."""
    def split(e, m, gamma_into_m):
        # p -> c becomes p -> m -> c; a gamma on p -> c (c is a reticulation) stays on the half into c
        p, c, g = edges[e]
        edges[e] = [p, m, gamma_into_m]
        edges.append([m, c, g])

    for _ in range(500):
        r = int(rng.integers(len(edges)))
        d = int(rng.integers(len(edges)))
        if r == d:
            continue
        pr, cr, _ = edges[r]
        pd, cd, _ = edges[d]
        lo_r, hi_r = nodes[cr][1], nodes[pr][1]
        if hi_r - lo_r < 1e-6:
            continue
        th = float(rng.uniform(lo_r + 0.1 * (hi_r - lo_r), hi_r - 0.1 * (hi_r - lo_r)))
        lo_d, hi_d = max(nodes[cd][1], th), nodes[pd][1]
        if hi_d - lo_d < 1e-6:
            continue
        tx = float(rng.uniform(lo_d + 0.05 * (hi_d - lo_d), hi_d - 0.05 * (hi_d - lo_d)))
        g = float(rng.uniform(*gamma_range))
        nodes.append(["#H%d" % k, th, True])
        h = len(nodes) - 1
        nodes.append(["x%d" % k, tx, False])
        x = len(nodes) - 1
        split(r, h, 1.0 - g)        # recipient edge: pr -> H -> cr
        split(d, x, None)           # donor edge:     pd -> X -> cd
        edges.append([x, h, g])     # hybrid edge X -> H
        return
    raise RuntimeError("could not place reticulation")


def enewick(nodes, edges):
    kids = {}
    for s, d, g in edges:
        kids.setdefault(s, []).append((d, g))
    has_parent = {d for _, d, _ in edges}
    root = [v for v in range(len(nodes)) if v not in has_parent][0]
    emitted = set()

    def rec(v, parent, gamma):
        label, h, is_ret = nodes[v]
        tail = ""
        if parent is not None:
            tail = ":" + repr(float(nodes[parent][1] - h))
            if gamma is not None:
                tail += "[&gamma=%r]" % float(gamma)
        if is_ret and v in emitted:
            return label + tail
        emitted.add(v)
        if v in kids:
            return "(" + ",".join(rec(c, v, g) for c, g in kids[v]) + ")" + label + tail
        return label + tail
    return rec(root, None, None) + ";"


# ------------------------------------------------------------------ flat exports
def export_network(net):
    idx, heights = build_network_msnc_index(net)
    gam = [idx.edge_gamma(e) for e in range(idx.n_edges)]
    return {
        "n_nodes": idx.n_nodes,
        "edge_src": list(idx.edge_src),
        "edge_dst": list(idx.edge_dst),
        "edge_gamma": [None if g is None else float(g) for g in gam],
        "is_retic": [bool(b) for b in idx.is_retic],
        "height": [float(h) for h in heights],
        "leaf_label": [idx.leaf_label.get(i) for i in range(idx.n_nodes)],
        "node_label": [n.label for n in idx._node_objs],
    }


def export_gene_tree(gt):
    nodes = list(gt.V())
    pos = {n: i for i, n in enumerate(nodes)}
    parent = [-1] * len(nodes)
    blen = [None] * len(nodes)
    for v in nodes:
        for c in gt.get_children(v):
            parent[pos[c]] = pos[v]
            e = gt.get_edge(v, c)
            if isinstance(e, list):
                e = e[0]
            ln = e.get_length()
            blen[pos[c]] = None if ln is None else float(ln)
    label = [n.label if not gt.get_children(n) else None for n in nodes]
    return {"parent": parent, "blen": blen, "label": label}


def timed_newick(flat, factor=1.0):
    """Newick with branch lengths (the reference's ``Network.newick()`` drops them)."""
    parent, blen, label = flat["parent"], flat["blen"], flat["label"]
    kids = [[] for _ in parent]
    root = 0
    for i, p in enumerate(parent):
        if p < 0:
            root = i
        else:
            kids[p].append(i)

    def rec(v):
        name = label[v] if label[v] else "g%d" % v
        s = "(" + ",".join(rec(c) for c in kids[v]) + ")" + name if kids[v] else name
        return s if parent[v] < 0 else s + ":" + repr(float(blen[v]) * factor)
    return rec(root) + ";"


def density(gt, net, species_of, theta, branch_thetas=None):
    v = gene_tree_msnc_log_density(gt, net, species_of, theta=theta, branch_thetas=branch_thetas)
    return None if (isinstance(v, float) and math.isinf(v)) else float(v)


def make_case(name, nw, alleles, theta, rng, n_trees, scales=(0.5, 0.9, 1.3), branch_thetas=None):
    net = Network.from_newick(nw, branch_length_unit=UNIT)
    species_of = {a: sp for sp, al in alleles.items() for a in al}
    trees, expect = [], []
    for k in range(n_trees):
        gt = simulate_gene_tree(net, alleles, theta, rng, branch_thetas=branch_thetas)
        flat = export_gene_tree(gt)
        trees.append(flat)
        expect.append(density(gt, net, species_of, theta, branch_thetas))
        for sc in (scales if k % 3 == 0 else ()):
            gtn = Network.from_newick(timed_newick(flat, sc), branch_length_unit=UNIT)
            trees.append(export_gene_tree(gtn))
            expect.append(density(gtn, net, species_of, theta, branch_thetas))
    case = {"name": name, "newick": nw, "network": export_network(net), "species_of": species_of,
            "theta": theta, "trees": trees, "expected": expect}
    if branch_thetas:   # JSON has no tuple keys: [[key, value], ...] with key a label or [parent, child]
        case["branch_thetas"] = [[list(k) if isinstance(k, tuple) else k, v] for k, v in branch_thetas.items()]
    return case


def main():
    rng = np.random.default_rng(4242)
    cases = []
    # the reference's own test networks (tests/test_mcmc_seq.py:224-227, :250-251)
    sp_g = ("((A:0.01,(B:0.005)#H1:0.005[&gamma=0.3]):0.01,(#H1:0.01[&gamma=0.7],C:0.015):0.005);")
    sp_1 = ("((A:0.01,(B:0.005)#H1:0.005[&gamma=1.0]):0.01,(#H1:0.01[&gamma=0.0],C:0.015):0.005);")
    one = {"A": ["A"], "B": ["B"], "C": ["C"]}
    cases.append(make_case("ref_test_gamma_0.3", sp_g, one, 0.02, rng, 6))
    cases.append(make_case("ref_test_gamma_1.0", sp_1, one, 0.02, rng, 4))
    bound = ("((C:0.9,((A:0.3,B:0.3)ab:0.3)#H1:0.3[&gamma=0.6])P1:0.2,(D:0.9,#H1:0.3[&gamma=0.4])P2:0.2)R;")
    cases.append(make_case("ref_test_bound_net", bound, {s: [s] for s in "ABCD"}, 0.5, rng, 6))
    cases.append(make_case("two_alleles_per_species", sp_g,
                           {"A": ["a1", "a2"], "B": ["b1", "b2", "b3"], "C": ["c1"]}, 0.02, rng, 6))
    # the fixed gene tree of TestReticulationDensity (only the A-side embedding is valid)
    for nm, nw in (("ref_gamma_mixture_g", sp_g), ("ref_gamma_mixture_1", sp_1)):
        net = Network.from_newick(nw, branch_length_unit=UNIT)
        gt = Network.from_newick("((A:0.012,B:0.012):0.013,C:0.025);", branch_length_unit=UNIT)
        so = {t: t for t in "ABC"}
        cases.append({"name": nm, "newick": nw, "network": export_network(net), "species_of": so, "theta": 0.02,
                      "trees": [export_gene_tree(gt)], "expected": [density(gt, net, so, 0.02)]})
    # random species trees and networks
    for n_leaves, n_retic, per_sp, theta, n_trees in [
            (2, 0, 1, 0.02, 4), (3, 0, 1, 0.02, 6), (5, 0, 1, 0.01, 6), (8, 0, 1, 0.02, 6), (16, 0, 1, 0.02, 5),
            (31, 0, 1, 0.03, 3), (6, 0, 2, 0.02, 5), (4, 0, 3, 0.05, 5),
            (4, 1, 1, 0.02, 8), (6, 1, 1, 0.02, 8), (6, 2, 1, 0.02, 8), (8, 2, 1, 0.01, 6), (8, 3, 1, 0.02, 6),
            (5, 2, 2, 0.02, 6), (10, 3, 1, 0.04, 5), (12, 4, 1, 0.02, 4),
            # more than 62 gene-tree nodes: the reference leaves its C path for Python integers (:58-63)
            (40, 0, 1, 0.02, 3), (20, 1, 2, 0.02, 3), (60, 0, 1, 0.02, 2)]:
        nodes, edges = random_network(n_leaves, n_retic, rng)
        nw = enewick(nodes, edges)
        alleles = {"s%d" % i: (["s%d" % i] if per_sp == 1 else ["s%d_%d" % (i, k) for k in range(per_sp)])
                   for i in range(n_leaves)}
        cases.append(make_case("rand_n%d_r%d_a%d" % (n_leaves, n_retic, per_sp), nw, alleles, theta, rng, n_trees))
    # fixed per-population thetas (resolve_branch_theta src/_units.py:90-125): edge keys, child keys, both root forms;
    # with them the reference takes its pure-Python DP (src/_msnc_density.py:229-233)
    cases.append(make_case("branch_thetas_ref_net", sp_g.replace("):0.01,(#H1", ")x:0.01,(#H1").replace("):0.005);", ")y:0.005)r;"),
                           one, 0.02, rng, 6, branch_thetas={("x", "A"): 0.01, "A": 0.5, "#H1": 0.03, ("y", "#H1"): 0.05,
                                                             "C": 0.04, ("__root__", "r"): 0.08}))
    nodes, edges = random_network(7, 2, rng)
    bt = {"i7": 0.05, ("i8", "i7"): 0.011, "s3": 0.004, "#H0": 0.07, "x1": 0.015, "__root__": 0.033, "s0": 0.02}
    cases.append(make_case("branch_thetas_rand_n7_r2", enewick(nodes, edges), {"s%d" % i: ["s%d" % i] for i in range(7)},
                           0.02, rng, 8, branch_thetas=bt))
    nodes, edges = random_network(9, 0, rng)
    root_label = nodes[-1][0]
    cases.append(make_case("branch_thetas_rand_n9_tree", enewick(nodes, edges), {"s%d" % i: ["s%d_a" % i, "s%d_b" % i] for i in range(9)},
                           0.03, rng, 6, branch_thetas={root_label: 0.2, "s1": 0.01, "s2": 0.02, "i9": 0.06, "i12": 0.002}))
    out = os.path.join(HERE, "msnc_cases.json")
    with open(out, "w") as f:
        json.dump({"source": "PhyNetPy v0.6.0 gene_tree_msnc_log_density; null = -inf", "cases": cases}, f)
    n = sum(len(c["trees"]) for c in cases)
    n_inf = sum(e is None for c in cases for e in c["expected"])
    print("wrote %s: %d cases, %d gene trees (%d with -inf)" % (out, len(cases), n, n_inf))


if __name__ == "__main__":
    main()
