"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference built into /tmp/ref, see
``_ref_loader.py``):

    python tests/golden/make_golden.py

Outputs (committed):
    felsen_cases.json    alignments + trees + models -> n_patterns, weights, pattern
                         keys, log-likelihood from FelsensteinCalculator
    pmatrix_cases.json   models x times -> eigen-system and P(t) from
                         _seq_likelihood.{JC69,HKY85,GTR}.p_matrix and GTR.py's expt
    engine_case.json     a small multi-locus Felsenstein-factor sum as computed by
                         the loop of _SeqLikelihoodEngine.log_likelihood

Trees are stored as flat arrays (parent / blen / label); the reference consumes
them through a duck-typed tree exposing exactly the calls FelsensteinCalculator
makes (root, get_children, get_edge(...).get_length(), node.label); for Newick
cases the same value is also computed through the reference's own
Network.from_newick and asserted equal, so the adapter is itself checked.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _ref_loader import load_reference  # noqa: E402

load_reference()
from phynetpy._seq_likelihood import (  # noqa: E402
    FelsensteinCalculator, SubstitutionModel, JC69, HKY85, GTR as SeqGTR, tip_partials_for_char,
)
from phynetpy.Network import Network  # noqa: E402
import phynetpy.GTR as RefGTR  # noqa: E402

from oracle.felsenstein_oracle import OracleTree  # noqa: E402  (only its Newick reader / flat form)


# ---------------------------------------------------------------- duck tree
class _N:
    def __init__(self, idx, label):
        self.idx = idx
        self.label = label


class _E:
    def __init__(self, length):
        self._l = length

    def get_length(self):
        return self._l


class DuckTree:
    def __init__(self, parent, blen, label):
        self.nodes = [_N(i, l) for i, l in enumerate(label)]
        self.kids = [[] for _ in parent]
        self._root = None
        for i, p in enumerate(parent):
            if p < 0:
                self._root = self.nodes[i]
            else:
                self.kids[p].append(self.nodes[i])
        self.blen = blen

    def root(self):
        return self._root

    def get_children(self, node):
        return self.kids[node.idx]

    def get_edge(self, node, child):
        return _E(self.blen[child.idx])


def make_model(spec):
    k = spec["kind"]
    if k == "JC69":
        return JC69()
    if k == "HKY85":
        return HKY85(spec["kappa"], spec["pi"])
    if k == "GTR":
        return SeqGTR(spec["pi"], spec["rates"])
    raise ValueError(k)


def run_case(name, aln_items, tree: OracleTree, model_spec, site_rate=1.0, newick=None):
    aln = dict(aln_items)
    fc = FelsensteinCalculator(aln)
    model = make_model(model_spec)
    duck = DuckTree(tree.parent, tree.blen, tree.label)
    ll = fc.log_likelihood(duck, model, site_rate=site_rate)
    if newick is not None:
        ll2 = fc.log_likelihood(Network.from_newick(newick), model, site_rate=site_rate)
        assert abs(ll - ll2) <= 1e-12 * max(1.0, abs(ll)), (name, ll, ll2)
    # pattern keys in first-occurrence order, reconstructed from the tip tables is lossy
    # (N and - share partials), so recompute keys the way the reference does (:349)
    taxa = list(aln.keys())
    seqs = [aln[t] for t in taxa]
    seen, keys = {}, []
    for col in range(len(seqs[0])):
        key = tuple(s[col].upper() for s in seqs)
        if key not in seen:
            seen[key] = len(keys)
            keys.append("\x1f".join(key))
    assert len(keys) == fc.n_patterns
    return {
        "name": name,
        "alignment": [[t, s] for t, s in aln_items],
        "tree": {"parent": tree.parent, "blen": tree.blen, "label": tree.label},
        "newick": newick,
        "model": model_spec,
        "site_rate": site_rate,
        "n_patterns": int(fc.n_patterns),
        "weights": [int(w) for w in fc._weights],
        "pattern_keys": keys,
        "logL": float(ll),
    }


def random_tree(rng, n_tips, labels, unary_prob=0.0, multif_prob=0.0, none_prob=0.0, zero_prob=0.0):
    """Random rooted tree by successive joins; optional polytomies / unary nodes / None lengths."""
    parent, blen, label = [], [], []

    def new(lab):
        parent.append(-1)
        blen.append(None)
        label.append(lab)
        return len(parent) - 1

    active = [new(l) for l in labels[:n_tips]]
    while len(active) > 1:
        k = 2
        if len(active) >= 3 and rng.random() < multif_prob:
            k = 3 if len(active) == 3 or rng.random() < 0.7 else 4
            k = min(k, len(active))
        picks = sorted(rng.choice(len(active), size=k, replace=False).tolist(), reverse=True)
        kids = [active.pop(i) for i in picks]
        p = new(None)
        for c in kids:
            parent[c] = p
        active.append(p)
        if rng.random() < unary_prob:
            u = new(None)
            parent[p] = u
            active[-1] = u
    for i in range(len(parent)):
        if parent[i] >= 0:
            r = rng.random()
            if r < none_prob:
                blen[i] = None
            elif r < none_prob + zero_prob:
                blen[i] = 0.0
            else:
                blen[i] = float(np.round(rng.exponential(0.08), 6))
    return OracleTree(parent, blen, label)


def random_alignment(rng, labels, L, alphabet):
    return [[lab, "".join(rng.choice(list(alphabet), size=L).tolist())] for lab in labels]


def main():
    cases = []
    jc = {"kind": "JC69"}
    hky = {"kind": "HKY85", "kappa": 2.5, "pi": [0.3, 0.2, 0.2, 0.3]}
    gtr_u = {"kind": "GTR", "pi": [0.25] * 4, "rates": [1, 2, 1, 1, 3, 1]}
    gtr_s = {"kind": "GTR", "pi": [0.1, 0.2, 0.3, 0.4], "rates": [1.0, 2.5, 0.7, 1.3, 3.1, 0.9]}

    # -- the reference's own fixtures (tests/test_mcmc_seq.py:108-109, tests/crosscheck/run_crosscheck.py:71-220)
    ALN = [["A", "ACGTACGTAC"], ["B", "ACGTACGAAC"], ["C", "ACGTTCGTAG"]]
    TREE = "((A:0.1,B:0.2):0.05,C:0.3):0.0;"
    for nm, ms in (("fixture_jc", jc), ("fixture_hky", hky), ("fixture_gtr_uniform", gtr_u), ("fixture_gtr_skewed", gtr_s)):
        cases.append(run_case(nm, ALN, OracleTree.from_newick(TREE), ms, newick=TREE))
    cc = [
        ("tree2taxa", "(A:0.08,B:0.08)g0;", jc,
         [["A", "ACGTACGTACGTACGTAACGTTGCA"], ["B", "ACGTACGAACGTAGGTAACGTTGCA"]]),
        ("tree3taxa_concordant", "((A:0.02,B:0.02)g1:0.03,C:0.05)g0;", jc,
         [["A", "ACGTACGTACGTACGTAACGTTGCAGGTACC"], ["B", "ACGTACGAACGTAGGTAACGTTGCAGGAACC"],
          ["C", "ACGTTCGAACGAAGGTAACCTTGCAGGTACT"]]),
        ("tree3taxa_deepILS", "((B:0.04,C:0.04)g1:0.02,A:0.06)g0;", jc,
         [["A", "ACGTACGTACGTACGTAACGTTGCAGGTACC"], ["B", "ACGTACGAACGTAGGTAACGTTGCAGGAACC"],
          ["C", "ACGTTCGAACGAAGGTAACCTTGCAGGTACT"]]),
        ("gtr_skewed", "((A:0.02,B:0.02)g1:0.03,C:0.05)g0;", gtr_s,
         [["A", "ACGTACGTACGTACGTAACGTTGCAGGTACCTTAGC"], ["B", "ACGTACGAACGTAGGTAACGTTGCAGGAACCTTAGG"],
          ["C", "ACGTTCGAACGAAGGTAACCTTGCAGGTACTTTAAC"]]),
        ("multiallele", "(((a1:0.005,a2:0.005)n1:0.015,b1:0.02)n2:0.02,c1:0.04)g0;", jc,
         [["a1", "ACGTACGTACGTACGTAACGTTGCAGGTACC"], ["a2", "ACGTACGTACGTACCTAACGTTGCAGGTACC"],
          ["b1", "ACGTACGAACGTAGGTAACGTTGCAGGAACC"], ["c1", "ACGTTCGAACGAAGGTAACCTTGCAGGTACT"]]),
        ("ambiguity_gaps", "((A:0.02,B:0.02)g1:0.03,C:0.05)g0;", jc,
         [["A", "ACGTRYSWKMBDHVN-ACGTACGT"], ["B", "ACGT-NACGTRYKMACGTAC-GTA"], ["C", "NNGTACGTACG--CGTRYACGTAC"]]),
        ("saturation", "((A:1.0,B:1.0)g1:1.5,C:2.5)g0;", jc,
         [["A", "ACGTACGTACGTACGTAACGTTGCAGGTACC"], ["B", "TGCATGCATGCATGCATTGCAACGTCCATGG"],
          ["C", "GGCCAATTGGCCAATTCCGGAATTGGCCAAT"]]),
        ("constant_sites", "((A:0.02,B:0.02)g1:0.03,C:0.05)g0;", jc,
         [["A", "AAAAAAAAAACCCCCCCCCC"], ["B", "AAAAAAAAAACCCCCCCCCC"], ["C", "AAAAAAAAAACCCCCCCCCC"]]),
        ("tree5taxa", "((((A:0.02,B:0.02)g1:0.01,C:0.03)g2:0.02,D:0.05)g3:0.02,E:0.07)g0;", jc,
         [["A", "ACGTACGTACGTACGTAACGTTGCAGGTACC"], ["B", "ACGTACGAACGTAGGTAACGTTGCAGGAACC"],
          ["C", "ACGTTCGAACGAAGGTAACCTTGCAGGTACT"], ["D", "ACGAACGTAGGTACGTAACGTAGCAGCTACC"],
          ["E", "ACGTACGTTCGTACGTATCGTTGGAGGTACA"]]),
    ]
    for nm, nw, ms, aln in cc:
        cases.append(run_case("crosscheck_" + nm, aln, OracleTree.from_newick(nw), ms, newick=nw))
    # ambiguity under a non-uniform model as well
    cases.append(run_case("ambiguity_gaps_gtr", cc[5][3], OracleTree.from_newick(cc[5][1]), gtr_s, newick=cc[5][1]))
    # polytomy through the reference's own Newick reader
    cases.append(run_case("polytomy_newick", ALN, OracleTree.from_newick("(A:0.1,B:0.2,C:0.3);"), gtr_s,
                          newick="(A:0.1,B:0.2,C:0.3);"))

    # -- randomised cases (seeded): shapes and edge cases the reference code path accepts
    rng = np.random.default_rng(20261019)
    plain, iupac = "ACGT", "ACGTACGTACGTACGTRYSWKMBDHVN-?X.OUacgtnZ*"
    models = [jc, hky, gtr_u, gtr_s,
              {"kind": "GTR", "pi": [0.4, 0.1, 0.15, 0.35], "rates": [0.3, 4.0, 0.9, 1.1, 6.0, 1.0]},
              {"kind": "HKY85", "kappa": 7.0, "pi": [0.2, 0.3, 0.3, 0.2]}]
    k = 0
    for n_tips, L, alpha, opts in [
        (4, 60, plain, {}), (5, 1000, plain, {}), (8, 200, plain, {}), (12, 300, plain, {}),
        (16, 500, plain, {}), (24, 400, plain, {}), (33, 150, plain, {}),
        (6, 120, iupac, {}), (9, 250, iupac, {}),
        (7, 150, plain, {"multif_prob": 0.6}), (10, 150, iupac, {"multif_prob": 0.4, "unary_prob": 0.3}),
        (8, 100, plain, {"unary_prob": 0.5}), (8, 100, plain, {"none_prob": 0.3, "zero_prob": 0.2}),
        (2, 40, plain, {}), (3, 1, plain, {}), (5, 700, "AC", {}),
    ]:
        for rep in range(2):
            labels = ["t%d" % i for i in range(n_tips)]
            tree = random_tree(rng, n_tips, labels, **opts)
            aln = random_alignment(rng, labels, L, alpha)
            ms = models[k % len(models)]
            sr = 1.0 if k % 3 else float(np.round(rng.uniform(0.3, 2.5), 3))
            cases.append(run_case("random_%02d_n%d_L%d" % (k, n_tips, L), aln, tree, ms, site_rate=sr))
            k += 1
    # leaves absent from the alignment (-> all-ones partials, :409-412) and extra alignment rows
    labels = ["t%d" % i for i in range(7)]
    tree = random_tree(rng, 7, labels)
    aln = random_alignment(rng, labels[:5] + ["unused"], 90, plain)
    cases.append(run_case("missing_taxa", aln, tree, gtr_s))
    # long branches / underflow clamp: 40 tips with branch length 40 -> site likelihood is pi-product, no clamp;
    # deep caterpillar with tiny probabilities exercising the 1e-300 floor
    labels = ["t%d" % i for i in range(260)]
    par, bl, lab = [], [], []
    prev = None
    for i, l in enumerate(labels):
        par.append(-1); bl.append(3.0); lab.append(l)
        tip = len(par) - 1
        if prev is None:
            prev = tip
        else:
            par.append(-1); bl.append(3.0); lab.append(None)
            node = len(par) - 1
            par[prev] = node; par[tip] = node
            prev = node
    bl[prev] = None
    deep = OracleTree(par, bl, lab)
    aln = random_alignment(rng, labels, 12, plain)
    cases.append(run_case("deep_caterpillar_260tips", aln, deep, gtr_s))
    # the 1e-300 clamp (:396): 48 tips joined by 1e-12 branches; a column with k substitutions has
    # likelihood ~1e-12k, so random columns underflow the floor while constant columns do not
    labels = ["t%d" % i for i in range(48)]
    tree = random_tree(rng, 48, labels)
    tree = OracleTree(tree.parent, [None if p < 0 else 1e-12 for p in tree.parent], tree.label)
    aln = random_alignment(rng, labels, 20, plain)
    aln = [[t, s + "AAAACCCC"] for t, s in aln]
    cases.append(run_case("underflow_clamp_tiny_branches", aln, tree, gtr_s))

    with open(os.path.join(HERE, "felsen_cases.json"), "w") as f:
        json.dump({"source": "PhyNetPy v0.6.0 FelsensteinCalculator (src/_seq_likelihood.py:311-426)", "cases": cases}, f)
    print("felsen cases:", len(cases))

    # ---------------------------------------------------------------- P(t)
    TIMES = [0.0, 1e-4, 1e-3, 0.01, 0.05, 0.1, 0.5, 1.0, 2.5, 10.0, 50.0, 500.0, -0.3]
    pm = []
    for ms in models:
        m = make_model(ms)
        pm.append({
            "family": "_seq_likelihood", "model": ms,
            "pi": m.pi.tolist(), "Q": m.Q.ravel().tolist(), "evals": m._evals.tolist(),
            "U": m._U.ravel().tolist(), "Uinv": m._Uinv.ravel().tolist(),
            "times": TIMES, "P": [m.p_matrix(t).ravel().tolist() for t in TIMES],
        })
    pub = [
        ("JC", [], {}), ("K80", [2.0], {}), ("F81", [[0.1, 0.2, 0.3, 0.4]], {}),
        ("HKY", [[0.3, 0.2, 0.2, 0.3], 2.5], {}), ("TN93", [[0.1, 0.2, 0.3, 0.4], 3.0, 5.0], {}),
        ("K81", [5.0, 1.0, 2.0], {}), ("SYM", [[1.0, 2.5, 0.7, 1.3, 3.1, 0.9]], {}),
        ("GTR", [[0.1, 0.2, 0.3, 0.4], [1.0, 2.5, 0.7, 1.3, 3.1, 0.9]], {}),
    ]
    for cls, args, _ in pub:
        m = getattr(RefGTR, cls)(*args)
        pm.append({
            "family": "GTR.py", "cls": cls, "args": args,
            "freqs": list(map(float, m.freqs)), "trans": list(map(float, m.trans)),
            "Q": np.asarray(m.getQ()).ravel().tolist(),
            "times": TIMES, "P": [np.asarray(m.expt(t)).ravel().tolist() for t in TIMES],
        })
    with open(os.path.join(HERE, "pmatrix_cases.json"), "w") as f:
        json.dump({"source": "PhyNetPy v0.6.0 p_matrix (src/_seq_likelihood.py:231-268) and expt (src/GTR.py)", "cases": pm}, f)
    print("pmatrix cases:", len(pm))

    # ------------------------------------------------- engine loop (Felsenstein factor)
    rng = np.random.default_rng(7)
    loci, per = [], []
    model = make_model(gtr_s)
    for i in range(12):
        n_tips = int(rng.integers(4, 11))
        labels = ["s%d" % j for j in range(n_tips)]
        tree = random_tree(rng, n_tips, labels)
        aln = random_alignment(rng, labels, int(rng.integers(80, 400)), plain)
        fc = FelsensteinCalculator(dict(aln))
        ll = fc.log_likelihood(DuckTree(tree.parent, tree.blen, tree.label), model)
        loci.append({"alignment": aln, "tree": {"parent": tree.parent, "blen": tree.blen, "label": tree.label}})
        per.append(float(ll))
    total = 0.0
    for v in per:  # src/_mcmc_seq.py:745-761 (Felsenstein terms only)
        total += v
    with open(os.path.join(HERE, "engine_case.json"), "w") as f:
        json.dump({"source": "loop of _SeqLikelihoodEngine.log_likelihood (src/_mcmc_seq.py:746-764), Felsenstein factor",
                   "model": gtr_s, "loci": loci, "per_locus": per, "total": total}, f)
    # tip encoding table for every byte value (src/_seq_likelihood.py:111-140)
    enc = {}
    for b in range(256):
        ch = chr(b)
        v = tip_partials_for_char(ch)
        enc[str(b)] = [int(x) for x in v]
    with open(os.path.join(HERE, "tip_encoding.json"), "w") as f:
        json.dump({"source": "tip_partials_for_char(chr(b)) for b in 0..255 (src/_seq_likelihood.py:120-140)", "table": enc}, f)
    print("done")


if __name__ == "__main__":
    main()
