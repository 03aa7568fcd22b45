"""GPU parity tests of K-m, the timed MSNC density log P(g | Psi), through the C ABI:
against the golden vectors frozen from the unmodified reference and against the CPU oracle on
seeded inputs.  Bar: 1e-10 relative (tree-shaped networks land within a few ulps); ``-inf`` exactly.
"""
import json
import math
import os

import numpy as np
import pytest

from oracle import msnc_oracle as MO

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "msnc_cases.json")))["cases"]
REL = 1e-10


def _pkg():
    from phynetpy_b200 import infer as P
    return P


def _close(got, want, rel=REL):
    if want is None or want == -math.inf:
        return got == -math.inf
    return abs(got - want) <= rel * max(1.0, abs(want))


def _thetas(case):
    bt = case.get("branch_thetas")
    return None if not bt else {(tuple(k) if isinstance(k, list) else k): v for k, v in bt}


def _trees(P, case):
    return [P.FlatTree(t["parent"], t["blen"], t["label"]) for t in case["trees"]]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_density(case):
    P = _pkg()
    net = P.SpeciesNetwork.from_dict(case["network"])
    got = P.msnc_log_density_batch(_trees(P, case), net, case["species_of"], case["theta"], branch_thetas=_thetas(case))
    assert got.shape == (len(case["trees"]),)
    for g, e in zip(got, case["expected"]):
        assert _close(float(g), e), (g, e)


def test_both_frontier_placements_agree(monkeypatch):
    """The frontier lives in shared memory when a block's worth fits, else in a global scratch buffer
    (PNB_MSNC_NO_SMEM=1 forces the latter): same DP source, identical results."""
    P = _pkg()
    ctx = P.default_context()
    for name in ("rand_n8_r3_a1", "rand_n16_r0_a1", "rand_n5_r2_a2", "rand_n20_r1_a2"):
        case = next(c for c in CASES if c["name"] == name)
        net = P.SpeciesNetwork.from_dict(case["network"])
        trees = _trees(P, case)
        monkeypatch.delenv("PNB_MSNC_NO_SMEM", raising=False)
        a = P.msnc_log_density_batch(trees, net, case["species_of"], case["theta"])
        smem_bytes = ctx.last_stats()["jobs_cached"]          # stats slot 6: bytes of shared memory per block
        monkeypatch.setenv("PNB_MSNC_NO_SMEM", "1")
        b = P.msnc_log_density_batch(trees, net, case["species_of"], case["theta"])
        assert ctx.last_stats()["jobs_cached"] == 0
        assert np.array_equal(a, b)
        if name == "rand_n16_r0_a1":
            assert smem_bytes > 0                              # a species tree always fits
    monkeypatch.delenv("PNB_MSNC_NO_SMEM", raising=False)


def test_warp_and_thread_kernels_agree(monkeypatch):
    """Reticulate networks run one WARP per gene tree (candidates of a step spread over the lanes, duplicates merged in
    candidate order); PNB_MSNC_SERIAL=1 runs the one-thread-per-tree DP instead.  Same entry order, same merge order,
    same floating-point operations: the two return the same bits, in shared memory and in the global scratch buffer."""
    P = _pkg()
    n_retic_cases = 0
    for case in CASES:
        net = P.SpeciesNetwork.from_dict(case["network"])
        if net.info(P.default_context())["reticulations"] == 0:
            continue
        n_retic_cases += 1
        trees = _trees(P, case)
        bt = _thetas(case)
        out = {}
        for serial in ("0", "1"):
            for no_smem in ("0", "1"):
                monkeypatch.setenv("PNB_MSNC_SERIAL", serial)
                monkeypatch.setenv("PNB_MSNC_NO_SMEM", no_smem)
                out[serial, no_smem] = P.msnc_log_density_batch(trees, net, case["species_of"], case["theta"], branch_thetas=bt)
        base = out["1", "0"]
        for key, v in out.items():
            assert np.array_equal(base, v, equal_nan=True), (case["name"], key)
    monkeypatch.delenv("PNB_MSNC_SERIAL", raising=False)
    monkeypatch.delenv("PNB_MSNC_NO_SMEM", raising=False)
    assert n_retic_cases >= 10


def test_species_tree_warp_kernel_agrees_with_thread_kernel(monkeypatch):
    """Species trees, small batches: one warp per gene tree, network nodes of one level side by side, factors added in
    step order afterwards (pnb_msnc_tree_warp_kernel).  PNB_MSNC_SERIAL=1 runs the one-thread-per-tree walk: same bits.
    Also on a 40-species tree (more network nodes than lanes) with several alleles per species."""
    from phynetpy_b200 import simulate as S
    P = _pkg()
    n_tree_cases = 0
    sets = []
    for case in CASES:
        net = P.SpeciesNetwork.from_dict(case["network"])
        if net.info(P.default_context())["reticulations"] == 0:
            sets.append((case["name"], net, _trees(P, case), case["species_of"], case["theta"], _thetas(case)))
    rng = np.random.default_rng(202)
    for n_species, alleles in ((40, 1), (12, 3), (3, 2)):
        net = S.random_species_network(n_species, 0, rng)
        trees, species_of = S.simulate_gene_trees_msnc(net, alleles, 0.02, rng, 25)
        sets.append(("sim_%d_%d" % (n_species, alleles), net, trees, species_of, 0.02, None))
    for name, net, trees, species_of, theta, bt in sets:
        n_tree_cases += 1
        monkeypatch.setenv("PNB_MSNC_SERIAL", "1")
        a = P.msnc_log_density_batch(trees, net, species_of, theta, branch_thetas=bt)
        monkeypatch.setenv("PNB_MSNC_SERIAL", "0")
        b = P.msnc_log_density_batch(trees, net, species_of, theta, branch_thetas=bt)
        assert np.array_equal(a, b, equal_nan=True), name
        if name.startswith("sim_"):
            d = net.to_dict()
            for t, g in list(zip(trees, b))[:5]:
                assert _close(float(g), MO.msnc_log_density(d, t.parent.tolist(), t.blen.tolist(), t.label, species_of, theta))
    monkeypatch.delenv("PNB_MSNC_SERIAL", raising=False)
    assert n_tree_cases >= 8


def test_warp_kernel_large_batch_with_duplicates(monkeypatch):
    """Many gene trees with several alleles per species under two / three reticulations: frontiers of hundreds of
    candidates per step with duplicate entries; blocks of four warps."""
    from phynetpy_b200 import simulate as S
    P = _pkg()
    rng = np.random.default_rng(77)
    for n_species, n_retic, alleles, n_trees in ((6, 2, 2, 1500), (8, 3, 1, 1500)):
        net = S.random_species_network(n_species, n_retic, rng)
        trees, species_of = S.simulate_gene_trees_msnc(net, alleles, 0.02, rng, 40)
        trees = (trees * (n_trees // len(trees) + 1))[:n_trees]
        monkeypatch.setenv("PNB_MSNC_SERIAL", "1")
        a = P.msnc_log_density_batch(trees, net, species_of, 0.02)
        monkeypatch.setenv("PNB_MSNC_SERIAL", "0")
        b = P.msnc_log_density_batch(trees, net, species_of, 0.02)
        assert np.array_equal(a, b)
        d = net.to_dict()
        for t, g in list(zip(trees, b))[:6]:
            assert _close(float(g), MO.msnc_log_density(d, t.parent.tolist(), t.blen.tolist(), t.label, species_of, 0.02))
    monkeypatch.delenv("PNB_MSNC_SERIAL", raising=False)


def test_global_scratch_is_used_in_several_launches_when_it_is_small(monkeypatch):
    """Frontiers that do not fit in shared memory go to a bounded global scratch buffer; a batch larger than the
    buffer is evaluated in several launches (PNB_MSNC_SCRATCH_MB shrinks the 2 GB default for this test)."""
    P = _pkg()
    ctx = P.default_context()
    case = next(c for c in CASES if c["name"] == "rand_n5_r2_a2")
    net = P.SpeciesNetwork.from_dict(case["network"])
    trees = _trees(P, case) * 60
    monkeypatch.delenv("PNB_MSNC_NO_SMEM", raising=False)
    monkeypatch.delenv("PNB_MSNC_SCRATCH_MB", raising=False)
    ref = P.msnc_log_density_batch(trees, net, case["species_of"], case["theta"])
    monkeypatch.setenv("PNB_MSNC_NO_SMEM", "1")
    monkeypatch.setenv("PNB_MSNC_SCRATCH_MB", "1")
    got = P.msnc_log_density_batch(trees, net, case["species_of"], case["theta"])
    assert ctx.last_stats()["launches"] > 1
    assert np.array_equal(ref, got)
    for g, e in zip(got[:len(case["expected"])], case["expected"]):
        assert _close(float(g), e)


def test_single_tree_entry_point_and_newick_network():
    P = _pkg()
    by = {c["name"]: c for c in CASES}
    case = by["ref_gamma_mixture_g"]
    gt = P.GeneTree.from_newick("((A:0.012,B:0.012):0.013,C:0.025);")
    so = {t: t for t in "ABC"}
    d_g = P.gene_tree_msnc_log_density(gt, case["newick"], so, theta=0.02)
    d_1 = P.gene_tree_msnc_log_density(gt, by["ref_gamma_mixture_1"]["newick"], so, theta=0.02)
    assert _close(d_g, case["expected"][0])
    assert d_g == pytest.approx(d_1 + math.log(0.3), abs=1e-9)      # tests/test_mcmc_seq.py:229-240
    with pytest.raises(ValueError, match="finite and positive"):
        P.gene_tree_msnc_log_density(gt, case["newick"], so, theta=0.0)
    # fixed per-population thetas: installed on the network handle, removed again by the next plain call
    bt_case = by["branch_thetas_ref_net"]
    net = P.SpeciesNetwork.from_newick(bt_case["newick"])
    t0 = bt_case["trees"][0]
    g0 = P.FlatTree(t0["parent"], t0["blen"], t0["label"])
    with_bt = P.gene_tree_msnc_log_density(g0, net, so, theta=0.02, branch_thetas=_thetas(bt_case))
    assert _close(with_bt, bt_case["expected"][0])
    plain = P.gene_tree_msnc_log_density(g0, net, so, theta=0.02)
    assert _close(plain, MO.msnc_log_density(bt_case["network"], t0["parent"], t0["blen"], t0["label"], so, 0.02))
    assert abs(plain - with_bt) > 1e-6
    with pytest.raises(ValueError, match="finite and positive"):
        P.gene_tree_msnc_log_density(g0, net, so, theta=0.02, branch_thetas={"A": -1.0})


def test_one_call_for_many_trees_equals_one_call_per_tree():
    P = _pkg()
    case = next(c for c in CASES if c["name"] == "rand_n8_r3_a1")
    net = P.SpeciesNetwork.from_dict(case["network"])
    trees = _trees(P, case)
    batch = P.msnc_log_density_batch(trees * 40, net, case["species_of"], case["theta"])   # > 256: parallel host prep
    solo = np.array([P.msnc_log_density_batch([t], net, case["species_of"], case["theta"])[0] for t in trees])
    assert np.array_equal(batch.reshape(40, -1), np.tile(solo, (40, 1)))


def test_mixed_tree_sizes_and_empty_tree():
    P = _pkg()
    case = next(c for c in CASES if c["name"] == "rand_n6_r1_a1")
    net = P.SpeciesNetwork.from_dict(case["network"])
    trees = _trees(P, case)[:3]
    # a gene tree over a subset of the species (pruned by hand: a cherry of two leaves)
    t0 = case["trees"][0]
    leaves = [i for i, l in enumerate(t0["label"]) if l is not None][:2]
    small = P.FlatTree([2, 2, -1], [0.2, 0.2, None], [t0["label"][leaves[0]], t0["label"][leaves[1]], None])
    empty = P.FlatTree([], [], [])
    got = P.msnc_log_density_batch([trees[0], small, empty, trees[1]], net, case["species_of"], case["theta"])
    want_small = MO.msnc_log_density(case["network"], [2, 2, -1], [0.2, 0.2, None], small.label, case["species_of"], case["theta"])
    assert _close(float(got[0]), case["expected"][0]) and _close(float(got[3]), case["expected"][1])
    assert _close(float(got[1]), want_small)
    assert got[2] == 0.0


def test_random_inputs_against_the_oracle():
    P = _pkg()
    rng = np.random.default_rng(2024)
    n_checked = 0
    for case in CASES:
        if case["name"].startswith("ref_gamma_mixture"):
            continue
        net = P.SpeciesNetwork.from_dict(case["network"])
        theta = case["theta"] * float(rng.uniform(0.5, 3.0))
        bt = _thetas(case)
        flats, want = [], []
        for t in case["trees"][:5]:
            bl = [None if b is None else b * float(rng.uniform(0.9, 1.5)) for b in t["blen"]]
            flats.append(P.FlatTree(t["parent"], bl, t["label"]))
            want.append(MO.msnc_log_density(case["network"], t["parent"], bl, t["label"], case["species_of"], theta, bt))
        got = P.msnc_log_density_batch(flats, net, case["species_of"], theta, branch_thetas=bt)
        for g, w in zip(got, want):
            assert _close(float(g), w), (case["name"], g, w)
            n_checked += 1
    assert n_checked > 80


def test_frontier_ceiling_is_reported_not_silently_wrong():
    from phynetpy_b200 import _lib
    P = _pkg()
    case = next(c for c in CASES if c["name"] == "rand_n5_r2_a2")
    net = P.SpeciesNetwork.from_dict(case["network"])
    ctx = P.default_context()
    _lib.check(_lib.load().pnb_msnc_net_set_max_frontier(net.handle(ctx), 1))
    with pytest.raises(P.PhyNetB200Error, match="frontier"):
        P.msnc_log_density_batch(_trees(P, case), net, case["species_of"], case["theta"])
    _lib.check(_lib.load().pnb_msnc_net_set_max_frontier(net.handle(ctx), 1 << 16))
    got = P.msnc_log_density_batch(_trees(P, case), net, case["species_of"], case["theta"])
    assert all(_close(float(g), e) for g, e in zip(got, case["expected"]))
    info = net.info(ctx)
    assert info["reticulations"] == 2 and info["species"] == 5


def test_engine_adds_the_coalescent_factor():
    """sum_i [log P(S_i | g_i) + log P(g_i | Psi)] with the reference's caching seam
    (src/_mcmc_seq.py:723-764): invalidate_locus / invalidate_network / clone / restore."""
    from phynetpy_b200 import simulate as S
    P = _pkg()
    case = next(c for c in CASES if c["name"] == "rand_n8_r2_a1")
    net = P.SpeciesNetwork.from_dict(case["network"])
    good = [i for i, e in enumerate(case["expected"]) if e is not None][:6]
    trees = [P.FlatTree(case["trees"][i]["parent"], case["trees"][i]["blen"], case["trees"][i]["label"]) for i in good]
    rng = np.random.default_rng(5)
    loci = []
    for ft in trees:   # the sequences only feed the Felsenstein factor, which has its own tests
        core = rng.integers(0, 4, size=120)
        loci.append({l: "".join("ACGT"[c] for c in np.where(rng.random(120) < 0.1, rng.integers(0, 4, size=120), core))
                     for l in ft.label if l is not None})
    model = P.GTR(S.DEFAULT_PI, S.DEFAULT_RATES)
    eng = P.SeqLikelihoodEngine(loci, case["species_of"], model)
    felsen_only = eng.log_likelihood(trees, None, 0.0)
    total = eng.log_likelihood(trees, net, case["theta"])
    want_msnc = [case["expected"][i] for i in good]
    assert total == pytest.approx(felsen_only + sum(want_msnc), rel=1e-12)
    for m, w in zip(eng._msnc, want_msnc):
        assert _close(m, w)
    # a theta move invalidates the network side only
    snap = eng.clone_caches()
    eng.invalidate_network()
    t2 = eng.log_likelihood(trees, net, case["theta"] * 2)
    w2 = [MO.msnc_log_density(case["network"], case["trees"][i]["parent"], case["trees"][i]["blen"], case["trees"][i]["label"],
                              case["species_of"], case["theta"] * 2) for i in good]
    assert t2 == pytest.approx(felsen_only + sum(w2), rel=1e-12)
    eng.restore_caches(snap)
    assert eng.log_likelihood(trees, net, case["theta"]) == total
    # a tree change followed by a network change: the packed store rewrites that locus' rows in place
    t3 = case["trees"][good[3]]
    stretched = P.FlatTree(t3["parent"], [None if b is None else b * 1.25 for b in t3["blen"]], t3["label"])
    new = list(trees)
    new[3] = stretched
    eng.invalidate_locus(3)
    eng.invalidate_network()
    eng.log_likelihood(new, net, case["theta"])
    w3 = MO.msnc_log_density(case["network"], t3["parent"], stretched.blen.tolist(), t3["label"], case["species_of"], case["theta"])
    assert _close(eng._msnc[3], w3)
    for k in (0, 1, 2, 4, 5):
        assert _close(eng._msnc[k], want_msnc[k])
    eng.invalidate_locus(3)
    assert eng.log_likelihood(trees, net, case["theta"]) == total
    # a gene tree that no longer embeds makes the state impossible
    bad_i = next(i for i, e in enumerate(case["expected"]) if e is None)
    bt = case["trees"][bad_i]
    if sorted(l for l in bt["label"] if l) == sorted(l for l in trees[0].label if l):
        new = list(trees)
        new[0] = P.FlatTree(bt["parent"], bt["blen"], bt["label"])
        eng.invalidate_locus(0)
        assert eng.log_likelihood(new, net, case["theta"]) == -math.inf
    # candidate sets: the MSNC half of _enumerate_scored_candidates in one call
    from phynetpy_b200 import candidates as CA
    arrs = CA.candidate_arrays(trees[1])
    res = eng.msnc_candidate_sets([(1, (trees[1], arrs[0], arrs[1])), (2, [trees[2], trees[2]])], net, case["theta"])
    assert res[0].shape[0] == arrs[0].shape[0] and res[1].shape[0] == 2
    for k in range(arrs[0].shape[0]):
        w = MO.msnc_log_density(case["network"], arrs[0][k].tolist(), arrs[1][k].tolist(), trees[1].label, case["species_of"],
                                case["theta"])
        assert _close(float(res[0][k]), w)
    assert _close(float(res[1][0]), want_msnc[2]) and res[1][0] == res[1][1]


def test_moved_networks_are_scored_like_fresh_ones():
    """Parameter moves of the species network (moves.net_slide_height_move / gamma_move) return a new flat
    network; its device handle is built on first use and the density follows the oracle."""
    from phynetpy_b200 import moves as MV, simulate as S
    P = _pkg()
    rng = np.random.default_rng(12)
    net = S.random_species_network(7, 2, rng)
    trees, species_of = S.simulate_gene_trees_msnc(net, 1, 0.02, rng, 12)
    cur, n_moved = net, 0
    for step in range(12):
        out = MV.net_slide_height_move(cur, rng) if step % 2 else MV.gamma_move(cur, rng)
        if out is None:
            continue
        cur = out[0]
        n_moved += 1
        got = P.msnc_log_density_batch(trees, cur, species_of, 0.02)
        d = cur.to_dict()
        for t, g in zip(trees, got):
            assert _close(float(g), MO.msnc_log_density(d, t.parent.tolist(), t.blen.tolist(), t.label, species_of, 0.02))
    assert n_moved >= 8
    base = P.msnc_log_density_batch(trees, net, species_of, 0.02)       # the original network is untouched
    d0 = net.to_dict()
    for t, g in zip(trees, base):
        assert _close(float(g), MO.msnc_log_density(d0, t.parent.tolist(), t.blen.tolist(), t.label, species_of, 0.02))
