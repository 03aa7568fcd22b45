"""CPU tests of the timed MSNC density path:

* the oracle (``oracle/msnc_oracle.py``) against the golden vectors produced by the unmodified
  reference (``tests/golden/msnc_cases.json``);
* the host logic of K-m -- network compiler, gene-tree preparation -- and the very DP source the CUDA
  kernel runs (``phynetpy_b200/csrc/pnb_msnc_core.hpp``), compiled here with g++ through
  ``tests/native/msnc_harness.cpp`` (test infrastructure, not part of the product library);
* the Python host side (``SpeciesNetwork``) and the C ABI on a host-only context.
"""
import ctypes as C
import json
import math
import os
import subprocess

import numpy as np
import pytest

from oracle import msnc_oracle as MO
from phynetpy_b200 import _lib
from phynetpy_b200._msnc_density import SpeciesNetwork, _check_theta
from phynetpy_b200.tree import FlatTree

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "msnc_cases.json")))["cases"]
REL = 1e-12   # the restatements follow the reference's operation order; bar is 1e-10


def _expected(e):
    return -math.inf if e is None else e


def _thetas(case):
    """branch_thetas of a golden case: JSON [[key, value], ...] -> {label | (parent, child): value}."""
    bt = case.get("branch_thetas")
    return None if not bt else {(tuple(k) if isinstance(k, list) else k): v for k, v in bt}


def _close(got, want):
    if want == -math.inf:
        return got == -math.inf
    return abs(got - want) <= REL * max(1.0, abs(want))


# ------------------------------------------------------------------ oracle vs the reference
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference(case):
    for t, e in zip(case["trees"], case["expected"]):
        got = MO.msnc_log_density(case["network"], t["parent"], t["blen"], t["label"], case["species_of"], case["theta"],
                                  _thetas(case))
        assert _close(got, _expected(e)), (got, e)


def test_golden_covers_the_interesting_regimes():
    names = {c["name"] for c in CASES}
    assert {"ref_gamma_mixture_g", "ref_gamma_mixture_1", "ref_test_bound_net"} <= names
    n_inf = sum(e is None for c in CASES for e in c["expected"])
    n_fin = sum(e is not None for c in CASES for e in c["expected"])
    assert n_inf >= 20 and n_fin >= 100
    assert any(sum(c["network"]["is_retic"]) >= 3 for c in CASES)
    assert any(max(len(t["parent"]) for t in c["trees"]) > 64 for c in CASES)      # two-word lineage sets
    assert any(len(set(c["species_of"].values())) < len(c["species_of"]) for c in CASES)  # several alleles per species
    with_thetas = [c for c in CASES if c.get("branch_thetas")]
    assert len(with_thetas) >= 3
    for c in with_thetas:   # the overrides matter: without them the density differs
        t = next(t for t, e in zip(c["trees"], c["expected"]) if e is not None)
        e = next(e for e in c["expected"] if e is not None)
        plain = MO.msnc_log_density(c["network"], t["parent"], t["blen"], t["label"], c["species_of"], c["theta"])
        assert abs(plain - e) > 1e-6


def test_reference_gamma_mixture_identity():
    # tests/test_mcmc_seq.py:229-240: only the A-side embedding is valid, so the density scales with gamma
    by = {c["name"]: c for c in CASES}
    d_g = by["ref_gamma_mixture_g"]["expected"][0]
    d_1 = by["ref_gamma_mixture_1"]["expected"][0]
    assert d_g == pytest.approx(d_1 + math.log(0.3), abs=1e-9)


# ------------------------------------------------------------------ the kernel's DP source on the host
@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("native") / "msnc_harness.so")
    src = os.path.join(HERE, "native", "msnc_harness.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True)
    lib = C.CDLL(out)
    lib.msnc_host_frontier_bound.restype = C.c_double
    return lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


def _pack(case, net=None):
    net = net or SpeciesNetwork.from_dict(case["network"])
    off, par, bl, sp = [0], [], [], []
    for t in case["trees"]:
        ft = FlatTree(t["parent"], t["blen"], t["label"])
        par.append(ft.parent)
        bl.append(ft.blen)
        sp.append(net.species_ids_of(ft, case["species_of"]))
        off.append(off[-1] + ft.n_nodes)
    return (net, np.array(off, np.int64), np.concatenate(par).astype(np.int32), np.concatenate(bl),
            np.concatenate(sp).astype(np.int32))


def _run_harness(lib, net, off, par, bl, sp, theta, cap=4096, force_nw=0, stride=1, branch_thetas=None, tree_path=1):
    n = len(off) - 1
    edge_theta, root_theta = net.branch_theta_arrays(branch_thetas)
    out = np.zeros(n)
    info = np.zeros(4, np.int32)
    err = C.create_string_buffer(256)
    rc = lib.msnc_host_eval(net.n_nodes, int(net.edge_src.shape[0]), _p(net.edge_src), _p(net.edge_dst), _p(net.edge_gamma),
                            _p(net.height), _p(net.node_species), len(net.species), C.c_int64(n), _p(off), _p(par), _p(bl),
                            _p(sp), C.c_double(theta), cap, force_nw, stride, _p(edge_theta) if edge_theta is not None else None,
                            C.c_double(root_theta), tree_path, _p(out), _p(info), err, 256)
    return rc, out, info, err.value.decode()


@pytest.mark.parametrize("force_nw,stride", [(0, 1), (2, 1), (0, 32)])
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_kernel_source_on_host_matches_reference(harness, case, force_nw, stride):
    # stride 32 = the interleaved layout the kernel uses when the frontier fits in shared memory
    net, off, par, bl, sp = _pack(case)
    rc, out, info, err = _run_harness(harness, net, off, par, bl, sp, case["theta"], force_nw=force_nw, stride=stride,
                                      cap=256 if stride > 1 else 4096, branch_thetas=_thetas(case))
    assert rc == 0, err
    assert info[2] == 0   # no frontier overflow
    for got, e in zip(out, case["expected"]):
        assert _close(got, _expected(e)), (got, e)


def test_tree_shaped_networks_are_bit_identical(harness):
    # no reticulation -> no log-sum-exp: the same additions in the same order as the reference
    for case in CASES:
        if any(case["network"]["is_retic"]):
            continue
        net, off, par, bl, sp = _pack(case)
        rc, out, info, err = _run_harness(harness, net, off, par, bl, sp, case["theta"], branch_thetas=_thetas(case))
        assert rc == 0, err
        # the visiting order differs from the reference's (depth-first, to keep entries narrow), which
        # permutes the per-branch terms of the sum: equal to a few ulps, not always to the bit
        for got, e in zip(out, case["expected"]):
            if e is None:
                assert got == -math.inf
            else:
                assert abs(got - e) <= 4 * np.spacing(abs(e)) * len(case["network"]["height"])


def test_tree_fast_path_equals_general_dp_bit_for_bit(harness):
    # species trees take msnc_dp_tree (one entry, updated in place); the general frontier DP must agree exactly
    n = 0
    for case in CASES:
        if any(case["network"]["is_retic"]):
            continue
        net, off, par, bl, sp = _pack(case)
        for nw, stride in ((0, 1), (2, 1), (0, 32)):
            a = _run_harness(harness, net, off, par, bl, sp, case["theta"], force_nw=nw, stride=stride, cap=64,
                             branch_thetas=_thetas(case), tree_path=1)
            b = _run_harness(harness, net, off, par, bl, sp, case["theta"], force_nw=nw, stride=stride, cap=64,
                             branch_thetas=_thetas(case), tree_path=0)
            assert a[0] == 0 and b[0] == 0
            assert np.array_equal(a[1], b[1])
            n += len(a[1])
    assert n > 300


def test_entry_width_stays_small(harness):
    # depth-first, widest-first ordering: the number of simultaneously open edges grows like the
    # depth of the network, not like the number of species (the reference's order opens one per species)
    for case in CASES:
        net, off, par, bl, sp = _pack(case)
        rc, out, info, err = _run_harness(harness, net, off[:2], par, bl, sp, case["theta"])
        assert rc == 0, err
        n_leaves = len(net.species)
        assert info[1] == net.n_nodes
        assert info[0] <= max(3, math.ceil(math.log2(max(n_leaves, 2))) + 2 + 2 * net.n_reticulations())


def test_frontier_overflow_is_reported(harness):
    case = next(c for c in CASES if c["name"] == "rand_n5_r2_a2")
    net, off, par, bl, sp = _pack(case)
    rc, out, info, err = _run_harness(harness, net, off, par, bl, sp, case["theta"], cap=2)
    assert rc == 0 and info[2] == 1
    assert np.isnan(out).any()


def test_frontier_bound(harness):
    case = next(c for c in CASES if c["name"] == "rand_n5_r2_a2")
    net = SpeciesNetwork.from_dict(case["network"])
    alleles = np.full(len(net.species), 2, np.int32)
    b = harness.msnc_host_frontier_bound(net.n_nodes, int(net.edge_src.shape[0]), _p(net.edge_src), _p(net.edge_dst),
                                         _p(net.edge_gamma), _p(net.height), _p(net.node_species), len(net.species), _p(alleles))
    assert b >= 4.0 and b == 2.0 ** round(math.log2(b))
    tree = next(c for c in CASES if c["name"] == "rand_n8_r0_a1")
    net = SpeciesNetwork.from_dict(tree["network"])
    alleles = np.ones(len(net.species), np.int32)
    b = harness.msnc_host_frontier_bound(net.n_nodes, int(net.edge_src.shape[0]), _p(net.edge_src), _p(net.edge_dst),
                                         _p(net.edge_gamma), _p(net.height), _p(net.node_species), len(net.species), _p(alleles))
    assert b == 1.0


def test_harness_agrees_with_oracle_on_fresh_random_inputs(harness):
    # inputs the golden file does not hold: perturbed heights, unmapped alleles, a missing species
    rng = np.random.default_rng(77)
    for case in CASES:
        if case["name"].startswith("ref_gamma_mixture"):
            continue
        bt = _thetas(case)
        netd = json.loads(json.dumps(case["network"]))
        species_of = dict(case["species_of"])
        drop = sorted(species_of)[0]
        del species_of[drop]            # this allele now maps nowhere: its lineage never enters
        sub = {"network": netd, "species_of": species_of, "trees": [], "theta": case["theta"] * 1.7}
        for t in case["trees"][:4]:
            bl = [None if b is None else b * float(rng.uniform(0.8, 1.6)) for b in t["blen"]]
            sub["trees"].append({"parent": t["parent"], "blen": bl, "label": t["label"]})
        net, off, par, bl, sp = _pack(sub)
        rc, out, info, err = _run_harness(harness, net, off, par, bl, sp, sub["theta"], branch_thetas=bt)
        assert rc == 0, err
        for t, got in zip(sub["trees"], out):
            want = MO.msnc_log_density(netd, t["parent"], t["blen"], t["label"], species_of, sub["theta"], bt)
            assert _close(got, want), (case["name"], got, want)


def test_harness_rejects_malformed_input(harness):
    case = CASES[0]
    net, off, par, bl, sp = _pack(case)
    bad = par.copy()
    bad[0] = 99
    rc, out, info, err = _run_harness(harness, net, off, bad, bl, sp, 0.02)
    assert rc == 1 and "out of range" in err
    cyc = par.copy()
    n0 = int(off[1])
    cyc[:n0] = (np.arange(n0) + 1) % n0     # a ring: no root
    rc, out, info, err = _run_harness(harness, net, off, cyc, bl, sp, 0.02)
    assert rc == 1 and ("no root" in err or "cycle" in err)
    two_roots = SpeciesNetwork([0, 1], [2, 3], [None, None], [1.0, 1.0, 0.0, 0.0], [None, None, "a", "b"])
    rc, out, info, err = _run_harness(harness, two_roots, off, par, bl, sp, 0.02)
    assert rc == 1 and "more than one root" in err


def test_empty_gene_tree_has_density_one(harness):
    case = CASES[0]
    net = SpeciesNetwork.from_dict(case["network"])
    off = np.array([0, 0], np.int64)
    z32, z64 = np.zeros(1, np.int32), np.zeros(1, np.float64)
    rc, out, info, err = _run_harness(harness, net, off, z32, z64, z32, 0.02)
    assert rc == 0 and out[0] == 0.0        # src/_msnc_density.py:222-223


# ------------------------------------------------------------------ Python host side
@pytest.mark.parametrize("case", [c for c in CASES if "newick" in c], ids=[c["name"] for c in CASES if "newick" in c])
def test_extended_newick_reader_rebuilds_the_reference_network(case):
    mine = SpeciesNetwork.from_newick(case["newick"])
    ref = SpeciesNetwork.from_dict(case["network"])
    assert mine.n_nodes == ref.n_nodes and mine.n_reticulations() == ref.n_reticulations()
    assert sorted(mine.species) == sorted(ref.species)
    # same density for the same gene trees (node / edge numbering is free)
    for t, e in list(zip(case["trees"], case["expected"]))[:5]:
        d = mine.to_dict()
        got = MO.msnc_log_density(d, t["parent"], t["blen"], t["label"], case["species_of"], case["theta"], _thetas(case))
        assert _close(got, _expected(e))


def test_from_network_duck_typing():
    class N:
        def __init__(self, label):
            self.label = label

    class E:
        def __init__(self, s, d, ln, g=None):
            self.src, self.dest, self._l, self._g = s, d, ln, g

        def get_length(self):
            return self._l

        def get_gamma(self):
            return self._g

    class Net:
        def __init__(self, nodes, edges):
            self._n, self._e = nodes, edges

        def V(self):
            return self._n

        def E(self):
            return self._e
    r, x, h, a, b, c = N("r"), N("x"), N("#H1"), N("A"), N("B"), N("C")
    y = N("y")
    net = Net([r, x, y, h, a, b, c], [E(r, x, 0.01), E(r, y, 0.005), E(x, a, 0.01), E(x, h, 0.005, 0.3), E(y, h, 0.01, 0.7),
                                      E(y, c, 0.015), E(h, b, 0.005)])
    sn = SpeciesNetwork.from_network(net)
    assert sn.species == ["A", "B", "C"]
    assert sn.height.tolist() == pytest.approx([0.02, 0.01, 0.015, 0.005, 0, 0, 0])
    assert sn.n_reticulations() == 1
    ft = FlatTree([2, 2, 4, 4, -1], [0.012, 0.012, 0.013, 0.025, None], ["A", "B", None, "C", None])
    assert sn.species_ids_of(ft, {"A": "A", "B": "B", "C": "C"}).tolist() == [0, 1, -1, 2, -1]
    assert sn.species_ids_of(ft, {"A": "A", "C": "nowhere"}).tolist() == [0, -1, -1, -1, -1]


def test_branch_theta_resolution_follows_the_reference():
    # resolve_branch_theta, src/_units.py:90-125: edge key before child key; root: ("__root__", label), "__root__", label
    sn = SpeciesNetwork.from_newick("((A:1,B:1)ab:1,C:2)r;")
    lab = sn.node_label
    e_of = {(lab[s], lab[d]): e for e, (s, d) in enumerate(zip(sn.edge_src.tolist(), sn.edge_dst.tolist()))}
    et, rt = sn.branch_theta_arrays({("ab", "A"): 0.1, "A": 0.5, "B": 0.2, "r": 0.9})
    assert et[e_of[("ab", "A")]] == 0.1 and et[e_of[("ab", "B")]] == 0.2
    assert np.isnan(et[e_of[("r", "C")]]) and np.isnan(et[e_of[("r", "ab")]])
    assert rt == 0.9
    assert sn.branch_theta_arrays({"__root__": 0.3, "r": 0.9})[1] == 0.3
    assert sn.branch_theta_arrays({("__root__", "r"): 0.4, "__root__": 0.3})[1] == 0.4
    none = sn.branch_theta_arrays(None)
    assert none[0] is None and np.isnan(none[1]) and sn.branch_theta_arrays({})[0] is None
    from phynetpy_b200._msnc_density import validate_branch_thetas
    with pytest.raises(ValueError, match="child labels or"):
        validate_branch_thetas({1: 0.1})
    with pytest.raises(ValueError, match="finite and positive"):
        validate_branch_thetas({"A": 0.0})
    with pytest.raises(ValueError, match="numeric"):
        validate_branch_thetas({"A": "x"})


def test_theta_validation_follows_the_reference():
    for bad in (0.0, -1.0, float("nan"), float("inf")):
        with pytest.raises(ValueError, match="finite and positive"):
            _check_theta(bad)
    with pytest.raises(ValueError, match="numeric"):
        _check_theta("x")
    assert _check_theta(0.5) == 0.5


# ------------------------------------------------------------------ C ABI without a device
def test_abi_compiles_networks_on_a_host_context_and_refuses_to_evaluate():
    lib = _lib.load()
    ctx = C.c_void_p()
    _lib.check(lib.pnb_ctx_create_host(C.byref(ctx)))
    try:
        for case in CASES[:8]:
            net = SpeciesNetwork.from_dict(case["network"])
            h = C.c_void_p()
            _lib.check(lib.pnb_msnc_net_create(ctx, net.n_nodes, int(net.edge_src.shape[0]), _lib.ptr(net.edge_src),
                                               _lib.ptr(net.edge_dst), _lib.ptr(net.edge_gamma), _lib.ptr(net.height),
                                               _lib.ptr(net.node_species), len(net.species), C.byref(h)))
            info = (C.c_int32 * 4)()
            _lib.check(lib.pnb_msnc_net_info(h, info))
            assert info[1] == net.n_nodes and info[2] == net.n_reticulations() and info[3] == len(net.species)
            out = np.zeros(1)
            off = np.array([0, 1], np.int64)
            par, bl, sp = np.array([-1], np.int32), np.zeros(1), np.array([0], np.int32)
            rc = lib.pnb_msnc_eval(ctx, h, 1, _lib.ptr(off), _lib.ptr(par), _lib.ptr(bl), _lib.ptr(sp), 0.02, _lib.ptr(out))
            assert rc == _lib.PNB_ERR_STATE
            assert b"no device" in lib.pnb_last_error()
            _lib.check(lib.pnb_msnc_net_destroy(h))
        # malformed network -> PNB_ERR_INVALID with the compiler's message
        h = C.c_void_p()
        src, dst = np.array([0, 1, 2], np.int32), np.array([1, 2, 0], np.int32)     # a ring
        gam, hh, sp = np.full(3, np.nan), np.zeros(3), np.full(3, -1, np.int32)
        rc = lib.pnb_msnc_net_create(ctx, 3, 3, _lib.ptr(src), _lib.ptr(dst), _lib.ptr(gam), _lib.ptr(hh), _lib.ptr(sp), 0,
                                     C.byref(h))
        assert rc == _lib.PNB_ERR_INVALID and b"no root" in lib.pnb_last_error()
    finally:
        lib.pnb_ctx_destroy(ctx)


def test_stress_own_generators_against_oracle(harness):
    """Networks and gene trees from the package's own generators (several alleles per species, up to four
    reticulations, frontiers of hundreds of entries with many merges): the kernel's DP source vs the oracle."""
    from phynetpy_b200 import simulate as S
    rng = np.random.default_rng(123)
    n_checked, max_frontier_seen = 0, 0
    for n_species, n_retic, alleles, n_trees in [(4, 1, 3, 6), (5, 2, 2, 6), (6, 3, 1, 8), (7, 4, 1, 6), (3, 2, 3, 6), (9, 0, 3, 4),
                                                 (6, 2, 2, 5)]:
        net = S.random_species_network(n_species, n_retic, rng)
        trees, species_of = S.simulate_gene_trees_msnc(net, alleles, 0.03, rng, n_trees)
        # half of the trees get perturbed heights: some stop embedding
        flats = []
        for k, t in enumerate(trees):
            flats.append(t if k % 2 == 0 else t.with_lengths(t.blen * float(rng.uniform(0.6, 1.4))))
        off = np.zeros(len(flats) + 1, np.int64)
        np.cumsum([t.n_nodes for t in flats], out=off[1:])
        par = np.concatenate([t.parent for t in flats]).astype(np.int32)
        bl = np.concatenate([t.blen for t in flats])
        sp = np.concatenate([net.species_ids_of(t, species_of) for t in flats]).astype(np.int32)
        rc, out, info, err = _run_harness(harness, net, off, par, bl, sp, 0.03, cap=1 << 14)
        assert rc == 0 and info[2] == 0, err
        d = net.to_dict()
        for t, got in zip(flats, out):
            want = MO.msnc_log_density(d, t.parent.tolist(), t.blen.tolist(), t.label, species_of, 0.03)
            assert _close(got, want), (n_species, n_retic, alleles, got, want)
            n_checked += 1
    assert n_checked >= 40


def test_tied_event_times_keep_node_order(harness):
    """Equal coalescence times (a zero-length internal branch): events are sorted stably, i.e. ties stay in node
    order -- the reference's ``events.sort(key=time)`` over its node order (src/_msnc_density.py:150-169)."""
    net = SpeciesNetwork.from_newick("((A:0.01,B:0.01)ab:0.02,C:0.03)r;")
    so = {"A": "A", "B": "B", "C": "C"}
    # node 3 = (A,B) at height 0.05, node 4 = root also at 0.05 (zero-length branch above node 3): child listed first
    child_first = FlatTree([3, 3, 4, 4, -1], [0.05, 0.05, 0.05, 0.0, None], ["A", "B", "C", None, None])
    # the same tree with the root listed BEFORE its child: the tie now puts the parent's event first
    parent_first = FlatTree([4, 4, 3, -1, 3], [0.05, 0.05, 0.05, None, 0.0], ["A", "B", "C", None, None])
    for ft in (child_first, parent_first):
        off = np.array([0, ft.n_nodes], np.int64)
        sp = net.species_ids_of(ft, so)
        rc, out, info, err = _run_harness(harness, net, off, ft.parent, ft.blen, sp, 0.02)
        assert rc == 0, err
        want = MO.msnc_log_density(net.to_dict(), ft.parent.tolist(), ft.blen.tolist(), ft.label, so, 0.02)
        assert _close(out[0], want)
    # with the parent's event first the two root children are not both present yet: no embedding
    off = np.array([0, 5], np.int64)
    a = _run_harness(harness, net, off, child_first.parent, child_first.blen, net.species_ids_of(child_first, so), 0.02)[1][0]
    b = _run_harness(harness, net, off, parent_first.parent, parent_first.blen, net.species_ids_of(parent_first, so), 0.02)[1][0]
    assert math.isfinite(a) and b == -math.inf
