"""Pin the CPU oracle to the reference: every golden vector frozen by
tests/golden/make_golden.py (outputs of the unmodified PhyNetPy v0.6.0) must be
reproduced by oracle/felsenstein_oracle.py.  Integer outputs bit-exactly,
floating point to 1e-12 relative (same NumPy algorithm, same operation order)."""
import math

import numpy as np
import pytest

from oracle import felsenstein_oracle as O
from conftest import load_golden

CASES = load_golden("felsen_cases.json")["cases"]
PM = load_golden("pmatrix_cases.json")["cases"]


def _model(spec):
    if spec["kind"] == "JC69":
        return O.jc69()
    if spec["kind"] == "HKY85":
        return O.hky85(spec["kappa"], spec["pi"])
    return O.gtr(spec["pi"], spec["rates"])


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_logL(case):
    locus = O.OracleLocus(dict(case["alignment"]))
    assert locus.n_patterns == case["n_patterns"]
    assert locus.int_weights.tolist() == case["weights"]
    assert ["\x1f".join(k) for k in locus.columns] == case["pattern_keys"]
    t = case["tree"]
    tree = O.OracleTree(t["parent"], t["blen"], t["label"])
    got = O.log_likelihood(locus, tree, _model(case["model"]), case["site_rate"])
    assert got == pytest.approx(case["logL"], rel=1e-12, abs=1e-12)


def test_oracle_codes_path_equals_partials_path():
    for case in CASES[:20]:
        locus = O.OracleLocus(dict(case["alignment"]))
        t = case["tree"]
        tree = O.OracleTree(t["parent"], t["blen"], t["label"])
        row = {name: i for i, name in enumerate(locus.taxa)}
        tip_row = [row.get(l, -1) if not tree.children[i] else -2 for i, l in enumerate(tree.label)]
        got = O.log_likelihood_codes(locus.codes(), locus.int_weights, tip_row, tree, _model(case["model"]), case["site_rate"])
        assert got == pytest.approx(case["logL"], rel=1e-12)


@pytest.mark.parametrize("case", [c for c in PM if c["family"] == "_seq_likelihood"],
                         ids=[str(c["model"]["kind"]) + str(i) for i, c in enumerate(PM) if c["family"] == "_seq_likelihood"])
def test_oracle_pmatrix_and_eigensystem(case):
    m = _model(case["model"])
    np.testing.assert_allclose(m.Q.ravel(), case["Q"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(m._evals, case["evals"], rtol=0, atol=1e-14)
    for t, P in zip(case["times"], case["P"]):
        np.testing.assert_allclose(m.p_matrix(t).ravel(), P, rtol=0, atol=1e-14)


def test_known_answer_survey_8c():
    # GTR([.1,.2,.3,.4],[1,2.5,.7,1.3,3.1,.9]) known answers
    m = O.gtr([0.1, 0.2, 0.3, 0.4], [1.0, 2.5, 0.7, 1.3, 3.1, 0.9])
    np.testing.assert_allclose(m._evals[:3], [-2.0733502804418866, -1.3692003983408958, -0.85726978800357212], atol=1e-13)
    np.testing.assert_allclose(m.p_matrix(0.1)[0], [0.8963030406252891, 0.01708872581189708, 0.06188053646830224,
                                                   0.0247276970945105], atol=1e-14)


def test_reference_known_answer_bruteforce():
    """The reference's own KAT (tests/test_mcmc_seq.py:107-121): pruning on compressed
    patterns == naive per-column pruning with scipy.linalg.expm, abs 1e-9."""
    expm = pytest.importorskip("scipy.linalg").expm
    aln = {"A": "ACGTACGTAC", "B": "ACGTACGAAC", "C": "ACGTTCGTAG"}
    tree = O.OracleTree.from_newick("((A:0.1,B:0.2):0.05,C:0.3):0.0;")
    for m in (O.jc69(), O.hky85(2.5, [0.3, 0.2, 0.2, 0.3]), O.gtr([0.25] * 4, [1, 2, 1, 1, 3, 1])):
        total = 0.0
        for col in range(10):
            def rec(node):
                kids = tree.children[node]
                if not kids:
                    return O.tip_partials_for_char(aln[tree.label[node]][col])
                r = np.ones(4)
                for c in kids:
                    r = r * (expm(m.Q * (tree.blen[c] or 0.0)) @ rec(c))
                return r
            total += math.log(float(m.pi @ rec(tree.root)))
        assert O.log_likelihood(O.OracleLocus(aln), tree, m) == pytest.approx(total, abs=1e-9)


def test_tip_encoding_table():
    table = load_golden("tip_encoding.json")["table"]
    for b in range(256):
        assert O.tip_partials_for_char(chr(b)).astype(int).tolist() == table[str(b)]


def test_compress_fast_equals_dict_compression():
    rng = np.random.default_rng(3)
    mat = rng.choice(np.frombuffer(b"ACGTN-", dtype=np.uint8), size=(6, 400))
    aln = {"t%d" % i: mat[i].tobytes().decode() for i in range(6)}
    locus = O.OracleLocus(aln)
    first, w, poc = O.compress_fast(mat)
    assert w.tolist() == locus.int_weights.tolist()
    assert poc.tolist() == locus.pattern_of_col.tolist()
    assert first.size == locus.n_patterns


def test_errors_match_reference():
    with pytest.raises(ValueError, match="Empty alignment"):
        O.OracleLocus({})
    with pytest.raises(ValueError, match="equally long"):
        O.OracleLocus({"a": "ACG", "b": "AC"})
    with pytest.raises(ValueError, match="sum to 1"):
        O.gtr([0.3, 0.3, 0.3, 0.3], [1] * 6)
