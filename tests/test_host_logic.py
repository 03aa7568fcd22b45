"""CPU: host-side logic -- tree marshalling, host pattern compression through the C ABI
(pure host code path of pnb_compress), alignment byte mapping, sharding, error behaviour."""
import math
import pickle

import numpy as np
import pytest

from oracle import felsenstein_oracle as O
from phynetpy_b200 import FlatTree, GeneTree, flatten, balanced_ranges, ShardPlan
from phynetpy_b200 import _seq_likelihood as SL
from phynetpy_b200 import simulate as S
from conftest import load_golden

CASES = load_golden("felsen_cases.json")["cases"]


def test_newick_roundtrip_matches_oracle_parser():
    for nw in ["((A:0.1,B:0.2):0.05,C:0.3):0.0;", "(A:0.1,B:0.2,C:0.3);",
               "((((A:0.02,B:0.02)g1:0.01,C:0.03)g2:0.02,D:0.05)g3:0.02,E:0.07)g0;"]:
        a = GeneTree.from_newick(nw).flat()
        b = O.OracleTree.from_newick(nw)
        assert a.parent.tolist() == b.parent
        assert a.label == b.label
        for i, (x, y) in enumerate(zip(a.blen, b.blen)):
            if b.parent[i] < 0:
                continue  # the root's own length is never read (reference :415-420)
            assert (math.isnan(x) and y is None) or x == y


def test_flatten_duck_typed_tree_preserves_structure():
    gt = GeneTree.from_newick("((A:0.1,B:0.2)x:0.05,(C:0.3,D)y:0.01)r;")

    class Duck:  # exposes only the calls the reference's Felsenstein path makes
        def root(self): return gt.root()
        def get_children(self, n): return gt.get_children(n)
        def get_edge(self, n, c): return [gt.get_edge(n, c)]  # list form (reference :417-418)

    ft = flatten(Duck())
    assert ft.n_nodes == 7
    assert sorted(l for l in ft.label if l in "ABCD") == list("ABCD")
    d = {ft.label[i]: ft.blen[i] for i in range(7) if ft.label[i]}
    assert d["A"] == 0.1 and d["C"] == 0.3 and math.isnan(d["D"]) and math.isnan(d["r"])
    rows = ft.tip_rows({"A": 0, "B": 1, "C": 2})
    assert sorted(rows.tolist()) == [-2, -2, -2, -1, 0, 1, 2]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_host_compress_bit_exact_vs_reference(case):
    """K-d host path through the C ABI: P, order and weights identical to the reference."""
    seqs = [s for _, s in case["alignment"]]
    mat, lut = SL._alignment_matrix(seqs)
    codes, w, first, poc = SL.compress_alignment(mat, lut, use_device=False)
    assert codes.shape[1] == case["n_patterns"]
    assert w.tolist() == case["weights"]
    keys = ["\x1f".join(s[c].upper() for s in seqs) for c in first]
    assert keys == case["pattern_keys"]
    locus = O.OracleLocus(dict(case["alignment"]))
    assert np.array_equal(codes, locus.codes())
    assert poc.tolist() == locus.pattern_of_col.tolist()


def test_compress_large_random_vs_oracle_fast():
    rng = np.random.default_rng(11)
    mat = S.ASCII_ACGT[rng.integers(0, 4, size=(7, 200000))]
    mat[:, 1000:1500] = mat[:, :500]  # guaranteed repeats
    codes, w, first, poc = SL.compress_alignment(mat, SL._CODE_OF_BYTE, use_device=False)
    f2, w2, poc2 = O.compress_fast(mat)
    assert np.array_equal(first, f2) and np.array_equal(w, w2) and np.array_equal(poc, poc2)
    assert int(w.sum()) == 200000


def test_non_ascii_characters_stay_distinct_patterns():
    mat, lut = SL._alignment_matrix(["AééN", "ACGT"])
    codes, w, first, _ = SL.compress_alignment(mat, lut, use_device=False)
    assert w.tolist() == [1, 1, 1, 1]
    assert codes[0].tolist() == [1, 15, 15, 15]


def test_constructor_errors_match_reference():
    with pytest.raises(ValueError, match="Empty alignment"):
        SL.FelsensteinCalculator({})
    with pytest.raises(ValueError, match="equally long"):
        SL.FelsensteinCalculator({"a": "ACG", "b": "AC"}, device_compress=False)
    with pytest.raises(ValueError, match="length 4"):
        SL.SubstitutionModel([0.5, 0.5], [1] * 6)
    with pytest.raises(ValueError, match="sum to 1"):
        SL.GTR([0.3] * 4, [1] * 6)


def test_model_eigensystem_matches_oracle_and_golden():
    pm = load_golden("pmatrix_cases.json")["cases"]
    for c in pm:
        if c["family"] != "_seq_likelihood":
            continue
        ms = c["model"]
        m = SL.JC69() if ms["kind"] == "JC69" else (SL.HKY85(ms["kappa"], ms["pi"]) if ms["kind"] == "HKY85" else SL.GTR(ms["pi"], ms["rates"]))
        np.testing.assert_allclose(m.Q.ravel(), c["Q"], atol=1e-15)
        np.testing.assert_allclose(m._evals, c["evals"], atol=1e-14)
        np.testing.assert_allclose(np.abs(m._U.ravel()), np.abs(c["U"]), atol=1e-12)


def test_device_kind_detection():
    class Custom(SL.SubstitutionModel):
        def p_matrix(self, t):
            return np.eye(4)
    assert SL._device_kind(SL.GTR([0.25] * 4, [1] * 6)) == "eigen"
    assert SL._device_kind(SL.HKY85(2.0, [0.25] * 4)) == "eigen"
    assert SL._device_kind(SL.JC69()) == "jc69"
    assert SL._device_kind(Custom([0.25] * 4, [1] * 6)) == "explicit"


def test_calculator_and_model_pickle_without_device_handles():
    fc = SL.FelsensteinCalculator({"A": "ACGTNN", "B": "ACGANN"}, device_compress=False)
    fc2 = pickle.loads(pickle.dumps(fc))
    assert fc2.taxa == ["A", "B"] and fc2.n_patterns == fc.n_patterns == 5
    assert fc2._locus is None
    m = pickle.loads(pickle.dumps(SL.GTR([0.1, 0.2, 0.3, 0.4], [1, 2, 3, 4, 5, 6])))
    assert "_pnb_handles" not in m.__dict__
    assert fc._weights.dtype == np.float64 and fc._weights.tolist() == [1.0, 1.0, 1.0, 1.0, 2.0]
    tp = fc._tip_partials
    assert tp["A"].shape == (5, 4) and tp["A"][0].tolist() == [1, 0, 0, 0]


def test_balanced_ranges_cover_and_balance():
    w = np.array([5, 1, 1, 1, 8, 2, 2, 4, 4, 4], dtype=float)
    for world in (1, 2, 3, 4, 8):
        r = balanced_ranges(w, world)
        assert r[0][0] == 0 and r[-1][1] == len(w)
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
    r = balanced_ranges(np.ones(10000), 8)
    assert [hi - lo for lo, hi in r] == [1250] * 8
    sp = ShardPlan(10, 1, 2, balanced_ranges(np.ones(10), 2))
    assert sp.local_range() == (5, 10) and sp.owner_of(3) == 0 and sp.owner_of(7) == 1


def test_simulator_shapes_and_tree_validity():
    rng = np.random.default_rng(5)
    par, bl = S.random_trees(4, 9, rng)
    assert par.shape == (4, 17) and (par[:, -1] == -1).all()
    assert (par[:, :-1] > np.arange(16)).all()
    assert np.nanmin(bl) > 0
    aln = S.simulate_alignments(par, bl, 300, S.DEFAULT_PI, S.DEFAULT_RATES, rng)
    assert aln.shape == (4, 9, 300) and set(np.unique(aln)) <= set(b"ACGT")
    # the oracle scores the simulated data (finite, negative)
    t = S.flat_trees(par, bl)[0]
    tree = O.OracleTree(t.parent.tolist(), [None if math.isnan(b) else float(b) for b in t.blen], t.label)
    ll = O.log_likelihood(O.OracleLocus(S.to_alignment_dict(aln[0])), tree, O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES))
    assert math.isfinite(ll) and ll < 0


def test_public_gtr_family_Q_matches_reference_golden():
    """phynetpy_b200.GTR (mirror of the reference's src/GTR.py): normalised Q, parameter
    mapping and validation, on the host; expt itself is checked on the GPU."""
    from phynetpy_b200 import GTR as G
    for c in load_golden("pmatrix_cases.json")["cases"]:
        if c["family"] != "GTR.py":
            continue
        m = getattr(G, c["cls"])(*c["args"])
        np.testing.assert_allclose(np.asarray(m.getQ()).ravel(), c["Q"], atol=1e-15)
        assert list(map(float, m.freqs)) == c["freqs"] and list(map(float, m.trans)) == c["trans"]
        assert m.state_count() == 4
        np.testing.assert_allclose(m.getQ().sum(axis=1), 0.0, atol=1e-15)
        assert -float(np.asarray(m.freqs) @ np.diag(m.getQ())) == pytest.approx(1.0, abs=1e-14)
    with pytest.raises(G.SubstitutionModelError):
        G.GTR([0.3, 0.3, 0.3, 0.3], [1] * 6)
    with pytest.raises(G.SubstitutionModelError):
        G.GTR([0.25] * 4, [1] * 5)
    with pytest.raises(G.SubstitutionModelError):
        G.K80(-1.0)
    with pytest.raises(G.SubstitutionModelError):
        G.JC().set_hyperparams({"kappa": 2.0})
    k = G.K81(5, 1, 2)
    k.set_hyperparams({"alpha": 9})
    assert k.trans == [1.0, 9.0, 2.0, 2.0, 9.0, 1.0]
    h = G.HKY([0.3, 0.2, 0.2, 0.3], 2.5)
    q0 = h.getQ().copy()
    h.set_hyperparams({"kappa": 4.0})
    assert h._decomp is None and not np.allclose(q0, h.getQ())
    assert G.K80(2.0).alpha == pytest.approx(0.5) and G.K80(2.0).beta == pytest.approx(0.25)


def test_peer_barrier_args_address_arithmetic():
    """ShardPlan.peer_barrier_args: the flag array sits behind the `copies` vectors of the symmetric allocation; a rank
    raises ITS slot in every peer's array and polls its own array at the peers' ranks."""
    import types
    from phynetpy_b200.distributed import ShardPlan
    world, per = 4, 250
    bases = [0x10000000 * (r + 1) for r in range(world)]
    hdl = types.SimpleNamespace(buffer_ptrs=bases)
    for rank in range(world):
        plan = ShardPlan(world * per, rank, world, [(r * per, (r + 1) * per) for r in range(world)])
        peer_flags, mine, peers = plan.peer_barrier_args(hdl, copies=2)
        off = 8 * 2 * per * world
        assert peers == [r for r in range(world) if r != rank]
        assert mine == bases[rank] + off
        assert peer_flags == [bases[r] + off + 8 * rank for r in peers]
        assert ShardPlan.BARRIER_SLOTS >= world
