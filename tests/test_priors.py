"""The network / theta prior on the flat species network against the reference's ``log_prior_seq``
(tests/golden/prior_cases.json: 11 networks x 6 hyper-parameter settings x 2 thetas)."""
import json
import math
import os

import pytest

from phynetpy_b200._msnc_density import SpeciesNetwork
from phynetpy_b200.priors import SeqPriors, log_prior_seq, reticulation_cycle_size

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "prior_cases.json")))["cases"]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_log_prior_equals_the_reference(case):
    for net in (SpeciesNetwork.from_dict(case["network"]), SpeciesNetwork.from_newick(case["newick"])):
        for rec in case["values"]:
            got = log_prior_seq(net, rec["theta"], SeqPriors(**rec["priors"]))
            if rec["value"] is None:
                assert got == -math.inf, (rec, got)
            else:
                assert got == pytest.approx(rec["value"], rel=1e-13, abs=1e-12), rec


def test_defaults_are_the_references():
    p = SeqPriors()
    assert (p.poisson_mean, p.max_reticulations, p.max_level, p.theta_shape, p.theta_prior_mean, p.theta_mean) == \
        (1.0, 4, None, 2.0, 0.036, 0.02)
    assert (p.diameter_mean, p.use_diameter_prior, p.gamma_alpha, p.gamma_beta, p.min_cycle_size, p.min_branch_length,
            p.branch_length_mean, p.use_branch_length_prior) == (0.5, True, 2.0, 2.0, 4, 1e-6, 0.1, True)


def test_cycle_sizes():
    # a reticulation between sister edges of a cherry's parent makes a 3-cycle; a bubble makes a 2-cycle
    parents = [[], [0], [0, 1], [2]]          # 0 root, 1 child of 0, 2 has parents 0 and 1 -> cycle {0, 1, 2}
    assert reticulation_cycle_size(parents, 2) == 3
    assert reticulation_cycle_size([[], [0, 0]], 1) == 2
    assert reticulation_cycle_size([[], [0]], 1) == 0
    assert log_prior_seq(SpeciesNetwork.from_newick("((A:1,B:1)x:1,C:2)r;"), 0.0, SeqPriors()) == -math.inf     # theta <= 0


def test_network_level_equals_the_reference():
    from phynetpy_b200.netmoves import network_level
    levels = set()
    for case in CASES:
        for net in (SpeciesNetwork.from_dict(case["network"]), SpeciesNetwork.from_newick(case["newick"])):
            assert network_level(net) == case["level"]
        levels.add(case["level"])
    assert {0, 1, 2, 3, 4} <= levels
    # several reticulations, yet level 1 (galled): the cap distinguishes blobs, not counts
    assert any(c["level"] == 1 and sum(c["network"]["is_retic"]) >= 2 for c in CASES)
