"""GPU: chains of the reference's gene-tree operators on flat arrays (``moves.slide_height_move`` = op_gene_node_height,
``moves.nni_move`` = op_gene_tree_nni, src/_mcmc_seq.py:1020-1113) through SeqLikelihoodEngine on the device:
after hundreds of accepted / rejected proposals with topology changes the cached per-locus values must equal a
from-scratch recomputation bit for bit, and the oracle to 1e-10 -- for the Felsenstein factor alone and for both
factors (K-b/K-c + K-m) on a network.  Also: device pattern compression == host compression at 10^6 columns."""
import math

import numpy as np
import pytest

from oracle import felsenstein_oracle as O, msnc_oracle as MO

pytestmark = pytest.mark.gpu
REL = 1e-10


def _otree(ft):
    return O.OracleTree(ft.parent.tolist(), [None if math.isnan(b) else float(b) for b in ft.blen], ft.label)


def _chain(eng, trees, net, theta, rng, n_steps, theta_moves):
    from phynetpy_b200 import moves as MV
    from phynetpy_b200.candidates import node_heights
    cur = list(trees)
    heights = [node_heights(t) for t in cur]
    total = eng.log_likelihood(cur, net, theta)
    acc = rej = topo = 0
    for step in range(n_steps):
        snap = eng.clone_caches()
        prop_h = None
        loghr = 0.0
        new_theta = theta
        if theta_moves and step % 10 == 9:
            new_theta = theta * math.exp(float(rng.uniform(-0.2, 0.2)))
            loghr = math.log(new_theta / theta)
            eng.invalidate_network()
            prop = cur
        else:
            i = int(rng.integers(len(cur)))
            if step % 3 == 2:
                o = MV.nni_move(cur[i], heights[i], rng)
                mv = None if o is None else (o[0], o[1], 0.0)
            else:
                mv = MV.slide_height_move(cur[i], heights[i], rng)
            if mv is None:
                continue
            topo += step % 3 == 2
            prop = list(cur)
            prop[i] = mv[0]
            prop_h = (i, mv[1])
            loghr = mv[2]
            eng.invalidate_locus(i)
        new_total = eng.log_likelihood(prop, net, new_theta)
        if math.log(rng.random()) < (new_total - total) + loghr:
            cur, total, theta = prop, new_total, new_theta
            if prop_h:
                heights[prop_h[0]] = prop_h[1]
            acc += 1
        else:
            eng.restore_caches(snap)
            rej += 1
    return cur, theta, total, acc, rej, topo


def test_flat_move_chain_felsenstein_cached_equals_from_scratch():
    from phynetpy_b200 import infer as P, simulate as S
    rng = np.random.default_rng(71)
    M, n, L = 60, 14, 400
    par, bl = S.random_trees(M, n, rng)
    aln = S.simulate_alignments(par, bl, L, S.DEFAULT_PI, S.DEFAULT_RATES, rng)
    trees = S.flat_trees(par, bl)
    data = [S.to_alignment_dict(aln[i]) for i in range(M)]
    model = P.GTR(S.DEFAULT_PI, S.DEFAULT_RATES)
    eng = P.SeqLikelihoodEngine(data, None, model)
    cur, _, total, acc, rej, topo = _chain(eng, trees, None, 0.0, rng, 500, theta_moves=False)
    assert acc > 50 and rej > 50 and topo > 50
    eng._commit_pending()
    cached = list(eng._felsen)
    for i in range(M):
        eng.invalidate_locus(i)
    eng.invalidate_device_cache()
    fresh = eng.log_likelihood(cur, None, 0.0)
    assert cached == list(eng._felsen) and fresh == total
    om = O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES)
    for i in range(0, M, 7):
        want = O.log_likelihood(O.OracleLocus(data[i]), _otree(cur[i]), om)
        assert abs(cached[i] - want) <= REL * abs(want)


def test_flat_move_chain_both_factors_on_a_network():
    from phynetpy_b200 import infer as P, simulate as S
    rng = np.random.default_rng(72)
    M, n_species, L = 40, 10, 300
    theta = 0.02
    net = S.random_species_network(n_species, 1, rng, height=0.005 * (n_species - 1))
    trees, species_of = S.simulate_gene_trees_msnc(net, 1, theta, rng, M)
    aln = S.simulate_alignments(np.stack([t.parent for t in trees]), np.stack([t.blen for t in trees]), L,
                                S.DEFAULT_PI, S.DEFAULT_RATES, rng)
    data = [S.to_alignment_dict(aln[i]) for i in range(M)]
    model = P.GTR(S.DEFAULT_PI, S.DEFAULT_RATES)
    eng = P.SeqLikelihoodEngine(data, species_of, model)
    cur, theta, total, acc, rej, topo = _chain(eng, trees, net, theta, rng, 400, theta_moves=True)
    assert acc > 20 and rej > 20
    eng._commit_pending()
    cf, cm = list(eng._felsen), list(eng._msnc)
    for i in range(M):
        eng.invalidate_locus(i)
    eng.invalidate_network()
    eng.invalidate_device_cache()
    fresh = eng.log_likelihood(cur, net, theta)
    assert cf == list(eng._felsen) and cm == list(eng._msnc) and fresh == total
    om = O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES)
    netd = net.to_dict()
    for i in range(0, M, 5):
        want = O.log_likelihood(O.OracleLocus(data[i]), _otree(cur[i]), om)
        assert abs(cf[i] - want) <= REL * abs(want)
        w2 = MO.msnc_log_density(netd, cur[i].parent.tolist(), cur[i].blen.tolist(), cur[i].label, species_of, theta)
        assert (cm[i] == w2) or abs(cm[i] - w2) <= REL * max(1.0, abs(w2))


def test_device_compression_equals_host_at_a_million_columns():
    """K-d at BASELINE cfg2 size: 50 taxa x 1.06e6 columns -- pattern count, first-occurrence order, weights,
    per-pattern tip codes and the column -> pattern map identical to the host implementation (which the small goldens
    pin to the reference's dict, src/_seq_likelihood.py:344-359)."""
    from phynetpy_b200 import simulate as S
    from phynetpy_b200._seq_likelihood import compress_alignment, _CODE_OF_BYTE
    rng = np.random.default_rng(2)
    par, bl = S.random_trees(1, 50, rng)
    blocks = [S.simulate_alignments(par, bl, 132_508, S.DEFAULT_PI, S.DEFAULT_RATES, rng)[0] for _ in range(8)]
    mat = np.concatenate(blocks, axis=1)
    # blemishes: gaps, ambiguity codes, repeated columns (duplicates far apart exercise the first-occurrence rule)
    mat[3, ::1001] = ord("-")
    mat[7, ::4999] = ord("N")
    mat[11, ::7001] = ord("R")
    mat[:, 900_000:900_500] = mat[:, 100:600]
    assert mat.shape == (50, 1_060_064)
    cd, wd, fd, pd_ = compress_alignment(mat, _CODE_OF_BYTE, use_device=True)
    ch, wh, fh, ph = compress_alignment(mat, _CODE_OF_BYTE, use_device=False)
    assert cd.shape == ch.shape and cd.shape[1] >= 1_000_000
    assert np.array_equal(wd, wh) and np.array_equal(fd, fh) and np.array_equal(pd_, ph) and np.array_equal(cd, ch)
    assert int(wd.sum()) == mat.shape[1] and bool((np.diff(fd) > 0).all())
    # the duplicated block maps onto the patterns of its first occurrence
    assert np.array_equal(pd_[900_000:900_500], pd_[100:600])


def test_device_sequence_simulation_equals_its_numpy_twin():
    """K-s (pnb_simulate_alignments; reference simulate_sequences / _evolve src/_sim_seq.py:431-504): same bytes as the
    NumPy twin for random binary trees, for a hand-numbered tree with a unary node and a polytomy, and independent of
    what else is in the batch."""
    from phynetpy_b200 import simulate as S
    from phynetpy_b200.tree import FlatTree
    rng = np.random.default_rng(9)
    par, bl = S.random_trees(40, 12, rng)
    trees = S.flat_trees(par, bl)
    odd = FlatTree([3, 3, 5, 5, 5, -1, 4], [0.1, 0.2, 0.05, 0.3, 0.0, float("nan"), 0.7], ["a", "b", "c", None, None, None, "d"])
    trees.append(odd)      # root 5 with children 2, 3, 4 (a polytomy); 4 is unary above leaf 6; leaves 0, 1, 2, 6
    ids = list(range(100, 100 + len(trees)))
    L = 1500
    dev = S.simulate_alignments_device(trees, L, S.DEFAULT_PI, S.DEFAULT_RATES, 77, ids)
    host = S.simulate_alignments_counter(trees, L, S.DEFAULT_PI, S.DEFAULT_RATES, 77, ids)
    assert len(dev) == len(host) == 41
    for d, h in zip(dev, host):
        assert d.shape == h.shape and np.array_equal(d, h)
    assert dev[-1].shape == (4, L)
    # the alignment of a tree depends on (seed, its stream id) only
    alone = S.simulate_alignments_device([trees[7]], L, S.DEFAULT_PI, S.DEFAULT_RATES, 77, [107])[0]
    assert np.array_equal(alone, dev[7])
    other_seed = S.simulate_alignments_device([trees[7]], L, S.DEFAULT_PI, S.DEFAULT_RATES, 78, [107])[0]
    assert not np.array_equal(other_seed, dev[7])
    # it is a GTR process: stationary frequencies at the tips
    freq = np.array([(np.concatenate([d.ravel() for d in dev[:40]]) == c).mean() for c in b"ACGT"])
    assert np.allclose(freq, S.DEFAULT_PI, atol=0.01)
