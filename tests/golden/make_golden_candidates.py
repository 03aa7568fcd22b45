"""Golden vectors for the candidate sets of the coupled moves, from the UNMODIFIED reference
(``_candidate_set`` / ``_gt_signature`` / ``_node_heights`` in src/_mcmc_seq.py, src/GraphUtils.py)
plus the Felsenstein log-likelihood of every candidate.  Build-container only (see _ref_loader.py).

    python tests/golden/make_golden_candidates.py   ->  tests/golden/candidate_cases.json
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
from phynetpy._mcmc_seq import _candidate_set, _gt_signature  # noqa: E402
from phynetpy.GraphUtils import _node_heights  # noqa: E402
from phynetpy._seq_likelihood import FelsensteinCalculator, GTR  # noqa: E402
from phynetpy.Network import Network  # noqa: E402
from phynetpy.models import BranchLengthUnit  # noqa: E402

from phynetpy_b200 import simulate as S  # noqa: E402  (input generator only)

PI, RATES = [0.1, 0.2, 0.3, 0.4], [1.0, 2.5, 0.7, 1.3, 3.1, 0.9]


def newick(parent, blen, label):
    kids = [[] for _ in parent]
    root = None
    for i, p in enumerate(parent):
        if p < 0:
            root = i
        else:
            kids[p].append(i)

    def rec(v):
        name = label[v] if label[v] else "n%d" % v
        s = "(" + ",".join(rec(c) for c in kids[v]) + ")" + name if kids[v] else name
        return s if parent[v] < 0 else s + ":" + repr(float(blen[v]))
    return rec(root) + ";"


def main():
    rng = np.random.default_rng(1900)
    model = GTR(PI, RATES)
    cases = []
    for n, L in [(4, 60), (6, 80), (9, 120), (13, 90)]:
        par, bl = S.random_trees(1, n, rng, height=0.12)
        aln = S.simulate_alignments(par, bl, L, PI, RATES, rng)[0]
        label = S.tip_labels(n) + [None] * (n - 1)
        nw = newick(par[0].tolist(), bl[0].tolist(), label)
        gt = Network.from_newick(nw, branch_length_unit=BranchLengthUnit.SUBSTITUTIONS_PER_SITE)
        heights = _node_heights(gt)
        cands = _candidate_set(gt, heights)
        fc = FelsensteinCalculator(S.to_alignment_dict(aln))
        out = []
        for cg, ch in cands:
            sig = _gt_signature(cg, ch)
            out.append({
                "signature": sorted([sorted(ls), h] for ls, h in sig),
                "logL": float(fc.log_likelihood(cg, model)),
            })
        cases.append({
            "tree": {"parent": par[0].tolist(), "blen": [None if np.isnan(b) else float(b) for b in bl[0]], "label": label},
            "alignment": [[k, v] for k, v in S.to_alignment_dict(aln).items()],
            "model": {"kind": "GTR", "pi": PI, "rates": RATES},
            "n_candidates": len(cands),
            "candidates": out,
        })
        print(n, len(cands))
    with open(os.path.join(HERE, "candidate_cases.json"), "w") as f:
        json.dump({"source": "PhyNetPy v0.6.0 _candidate_set (src/_mcmc_seq.py:1836-1861) + FelsensteinCalculator", "cases": cases}, f)


if __name__ == "__main__":
    main()
