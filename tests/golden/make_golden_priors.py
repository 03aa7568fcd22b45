"""Golden vectors for the network / theta prior, from the UNMODIFIED reference (``log_prior_seq``
``src/_mcmc_seq.py:294-411``).  Build-container only.

    python tests/golden/make_golden_priors.py   ->  tests/golden/prior_cases.json
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
from phynetpy._mcmc_seq import MCMCSeqPriors, log_prior_seq  # noqa: E402
from phynetpy.GraphUtils import level  # noqa: E402
from phynetpy.Network import Network  # noqa: E402
from phynetpy.models import BranchLengthUnit  # noqa: E402

from make_golden_msnc import enewick, export_network, random_network  # noqa: E402

UNIT = BranchLengthUnit.SUBSTITUTIONS_PER_SITE

SETTINGS = [
    {},                                                            # the reference's defaults
    {"gamma_alpha": 1.0, "gamma_beta": 1.0, "use_diameter_prior": False},
    {"min_cycle_size": 0, "min_branch_length": 0.0},
    {"min_cycle_size": 5, "poisson_mean": 2.5, "theta_shape": 3.0, "theta_prior_mean": 0.01},
    {"use_branch_length_prior": False, "diameter_mean": 0.02, "max_reticulations": 2},
    {"min_branch_length": 0.004, "branch_length_mean": 0.02},
]


def main():
    rng = np.random.default_rng(99)
    cases = []
    nets = [(3, 0), (5, 0), (4, 1), (5, 1), (6, 2), (7, 3), (6, 1), (8, 4), (5, 2), (4, 2), (12, 3), (16, 4), (10, 5)]
    for n_leaves, n_retic in nets:
        nodes, edges = random_network(n_leaves, n_retic, rng, height=0.06)
        nw = enewick(nodes, edges)
        net = Network.from_newick(nw, branch_length_unit=UNIT)
        vals = []
        for kw in SETTINGS:
            pri = MCMCSeqPriors(**kw)
            for theta in (0.02, 0.003):
                v = log_prior_seq(net, theta, pri)
                vals.append({"priors": kw, "theta": theta, "value": None if math.isinf(v) else float(v)})
        cases.append({"name": "n%d_r%d" % (n_leaves, n_retic), "newick": nw, "network": export_network(net), "values": vals,
                      "level": int(level(net))})
        print(cases[-1]["name"], [None if v["value"] is None else round(v["value"], 3) for v in vals[:6]])
    # the reference's own bubble / tiny-cycle guard: a reticulation whose parents are parent and child (3-cycle)
    nw = "((A:0.02,(B:0.01)#H1:0.01[&gamma=0.4])x:0.01,#H1:0.02[&gamma=0.6])r;"
    net = Network.from_newick(nw, branch_length_unit=UNIT)
    vals = []
    for kw in SETTINGS:
        v = log_prior_seq(net, 0.02, MCMCSeqPriors(**kw))
        vals.append({"priors": kw, "theta": 0.02, "value": None if math.isinf(v) else float(v)})
    cases.append({"name": "three_cycle", "newick": nw, "network": export_network(net), "values": vals, "level": int(level(net))})
    print(cases[-1]["name"], [v["value"] for v in vals])
    out = os.path.join(HERE, "prior_cases.json")
    with open(out, "w") as f:
        json.dump({"source": "PhyNetPy v0.6.0 log_prior_seq", "cases": cases}, f)
    print("wrote", out)


if __name__ == "__main__":
    main()
