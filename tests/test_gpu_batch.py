"""GPU: resident batches (``pnb_batch_*``) -- programs stay in HBM, evaluations send branch lengths only.
Bars: bit-identical to ``pnb_eval(PNB_EVAL_FULL)`` on the same trees for every input route (pageable host
lengths, the batch's pinned buffer, caller-pinned memory, device lengths, device output on the caller's stream,
one chunk and several), the oracle to 1e-10, and the loci's partial caches describe the last evaluation (a
dirty-path ``pnb_eval`` afterwards recomputes only the changed path and equals a full recomputation)."""
import math

import numpy as np
import pytest

from oracle import felsenstein_oracle as O

pytestmark = pytest.mark.gpu
REL = 1e-10


def _world(M, n, L, seed, ragged=False):
    from phynetpy_b200 import infer as P, simulate as S, _lib
    from phynetpy_b200._seq_likelihood import _model_handle
    from phynetpy_b200.device import pack_trees
    rng = np.random.default_rng(seed)
    par, bl = S.random_trees(M, n, rng)
    aln = S.simulate_alignments(par, bl, L, S.DEFAULT_PI, S.DEFAULT_RATES, rng)
    trees = S.flat_trees(par, bl)
    taxa = S.tip_labels(n)
    calcs = []
    for i in range(M):
        a = aln[i][:, : (L - (i * 37) % (L // 2))] if ragged else aln[i]
        calcs.append(P.FelsensteinCalculator.from_matrix(taxa, np.ascontiguousarray(a), device_compress=False))
    model = P.GTR(S.DEFAULT_PI, S.DEFAULT_RATES)
    ctx = calcs[0]._context()
    for c in calcs:
        c._locus_handle()
    handles = np.array([c._locus.value for c in calcs], dtype=np.uint64)
    node_off, parent, blen, tip_row = pack_trees(trees, [c._row_of for c in calcs])
    mh = _model_handle(model, ctx)
    return dict(P=P, S=S, lib=_lib, ctx=ctx, calcs=calcs, trees=trees, model=model, mh=mh, handles=handles,
                node_off=node_off, parent=parent, blen=blen, tip_row=tip_row, aln=aln)


def _otree(ft):
    return O.OracleTree(ft.parent.tolist(), [None if math.isnan(b) else float(b) for b in ft.blen], ft.label)


def test_batch_equals_pnb_eval_for_every_input_route():
    import torch
    from phynetpy_b200.device import ResidentBatch
    w = _world(M=300, n=12, L=400, seed=7, ragged=True)      # PNB_CHUNK_MIN_JOBS=128 (conftest): several upload chunks
    ctx, lib = w["ctx"], w["lib"]
    b = ResidentBatch(ctx, w["mh"], w["handles"], w["node_off"], w["parent"], w["tip_row"])
    assert b.info["jobs"] == 300 and b.info["nodes"] == w["parent"].shape[0] and b.info["chunks"] >= 2
    om = O.gtr(w["S"].DEFAULT_PI, w["S"].DEFAULT_RATES)
    for k in range(3):
        bl = np.ascontiguousarray(w["blen"] * (1.0 + 0.05 * k))
        want = ctx.eval(w["mh"], w["handles"], w["node_off"], w["parent"], bl, w["tip_row"], 1.0, lib.EVAL_FULL)
        got_pageable = b.eval(w["mh"], bl)
        assert np.array_equal(got_pageable, want)
        assert ctx.last_stats()["updates"] == b.info["updates"] and ctx.last_stats()["h2d_bytes"] == 8 * b.info["nodes"]
        b.blen_host[:] = bl
        assert np.array_equal(b.eval(w["mh"], b.blen_host), want)
        pinned = torch.empty(bl.shape[0], dtype=torch.float64).pin_memory()
        pinned.numpy()[:] = bl
        assert np.array_equal(b.eval(w["mh"], pinned.numpy()), want)
        d_bl = torch.from_numpy(bl).cuda()
        assert np.array_equal(b.eval(w["mh"], d_bl.data_ptr()), want)
        assert ctx.last_stats()["h2d_bytes"] == 0 and ctx.last_stats()["launches"] == 2
        # device lengths in, device values out, on the caller's stream
        stream = torch.cuda.Stream()
        ctx.set_stream(stream.cuda_stream)
        try:
            out = torch.full((302,), -1.0, dtype=torch.float64, device="cuda")
            with torch.cuda.stream(stream):
                b.eval(w["mh"], d_bl.data_ptr(), out_device_ptr=out.data_ptr() + 8)
                twice = out * 2.0
            stream.synchronize()
            assert np.array_equal(out.cpu().numpy()[1:-1], want) and out[0].item() == -1.0 and out[-1].item() == -1.0
            assert np.array_equal(twice.cpu().numpy()[1:-1], 2.0 * want)
        finally:
            ctx.set_stream(None)
    # the oracle on a sample of the last evaluation
    for i in (0, 151, 299):
        t = w["trees"][i].with_lengths(w["trees"][i].blen * 1.1)
        exp = O.log_likelihood_codes(w["calcs"][i]._codes, w["calcs"][i]._int_weights,
                                     t.tip_rows(w["calcs"][i]._row_of).tolist(), _otree(t), om)
        assert abs(want[i] - exp) <= REL * abs(exp)
    b.close()


def test_loci_describe_the_last_batch_evaluation():
    """After a batch evaluation (committed mode) a one-branch change on one locus is a dirty-path evaluation:
    few nodes recomputed, value equal to a full recomputation bit for bit -- for host and for device lengths."""
    import torch
    from phynetpy_b200.device import ResidentBatch
    w = _world(M=40, n=16, L=300, seed=11)
    ctx, lib, calcs, trees, model = w["ctx"], w["lib"], w["calcs"], w["trees"], w["model"]
    b = ResidentBatch(ctx, w["mh"], w["handles"], w["node_off"], w["parent"], w["tip_row"])
    for route in ("host", "device"):
        bl = np.ascontiguousarray(w["blen"] * (1.3 if route == "host" else 0.8))
        if route == "host":
            vals = b.eval(w["mh"], bl)
        else:
            d_bl = torch.from_numpy(bl).cuda()
            vals = b.eval(w["mh"], d_bl.data_ptr())
        for i in (3, 17, 39):
            a, e = int(w["node_off"][i]), int(w["node_off"][i + 1])
            cur = trees[i].with_lengths(bl[a:e].copy())
            # the unchanged tree is served from the cache (root clean): no device work, same bits
            assert calcs[i].log_likelihood(cur, model) == vals[i]
            assert ctx.last_stats()["nodes_recomputed"] == 0
            bl2 = bl[a:e].copy()
            leaf = int(np.nonzero(trees[i].n_children() == 0)[0][0])
            bl2[leaf] *= 1.7
            moved = trees[i].with_lengths(bl2)
            v = calcs[i].log_likelihood(moved, model)
            st = ctx.last_stats()
            assert 0 < st["nodes_recomputed"] < 15
            assert v == calcs[i].log_likelihood(moved, model, full=True)
        # the batch still evaluates (it rewrites the slots it owns) and still equals pnb_eval
        want = ctx.eval(w["mh"], w["handles"], w["node_off"], w["parent"], bl, w["tip_row"], 1.0, lib.EVAL_FULL)
        assert np.array_equal(b.eval(w["mh"], bl), want) and np.array_equal(want, vals)
    b.close()
    # after the batch is gone the loci keep what it last computed
    i = 5
    a, e = int(w["node_off"][i]), int(w["node_off"][i + 1])
    assert calcs[i].log_likelihood(trees[i].with_lengths(bl[a:e].copy()), model) == vals[i]
    assert ctx.last_stats()["nodes_recomputed"] == 0


def test_score_only_batch_model_change_and_rescale():
    from phynetpy_b200.device import ResidentBatch
    from phynetpy_b200._seq_likelihood import _model_handle
    w = _world(M=24, n=20, L=250, seed=13)
    ctx, lib, P = w["ctx"], w["lib"], w["P"]
    want = ctx.eval(w["mh"], w["handles"], w["node_off"], w["parent"], w["blen"], w["tip_row"], 1.0, lib.EVAL_FULL)
    nb = ResidentBatch(ctx, w["mh"], w["handles"], w["node_off"], w["parent"], w["tip_row"], lib.EVAL_NOSTORE)
    assert np.array_equal(nb.eval(w["mh"], w["blen"]), want)
    # a score-only batch may hold a locus several times
    twice = ResidentBatch(ctx, w["mh"], np.concatenate([w["handles"][:3], w["handles"][:3]]),
                          np.concatenate([w["node_off"][:4], w["node_off"][1:4] + w["node_off"][3]]),
                          np.concatenate([w["parent"][: w["node_off"][3]]] * 2),
                          np.concatenate([w["tip_row"][: w["node_off"][3]]] * 2), lib.EVAL_NOSTORE)
    v = twice.eval(w["mh"], np.concatenate([w["blen"][: w["node_off"][3]]] * 2))
    assert np.array_equal(v[:3], want[:3]) and np.array_equal(v[3:], want[:3])
    # another model through the same programs (site_rate too)
    hky = P.HKY85(2.5, [0.3, 0.2, 0.2, 0.3])
    mh2 = _model_handle(hky, ctx)
    want2 = ctx.eval(mh2, w["handles"], w["node_off"], w["parent"], w["blen"], w["tip_row"], 0.7, lib.EVAL_FULL)
    assert np.array_equal(nb.eval(mh2, w["blen"], site_rate=0.7), want2)
    sb = ResidentBatch(ctx, w["mh"], w["handles"], w["node_off"], w["parent"], w["tip_row"])
    assert np.array_equal(sb.eval(mh2, w["blen"], site_rate=0.7), want2)
    # the loci's caches carry the model of the last batch evaluation: a call with the first model recomputes all
    assert w["calcs"][2].log_likelihood(w["trees"][2], w["model"]) == want[2]
    assert ctx.last_stats()["nodes_recomputed"] == 19
    # exact rescaling mode
    rb = ResidentBatch(ctx, w["mh"], w["handles"], w["node_off"], w["parent"], w["tip_row"], lib.EVAL_RESCALE)
    want_r = ctx.eval(w["mh"], w["handles"], w["node_off"], w["parent"], w["blen"], w["tip_row"], 1.0,
                      lib.EVAL_FULL | lib.EVAL_RESCALE)
    got_r = rb.eval(w["mh"], w["blen"])
    assert np.array_equal(got_r, want_r) and np.allclose(got_r, want, rtol=1e-12)
    for x in (nb, twice, sb, rb):
        x.close()


def test_batch_refuses_pending_loci_and_duplicates_and_dies_with_its_loci():
    from phynetpy_b200 import PhyNetB200Error
    from phynetpy_b200.device import ResidentBatch
    w = _world(M=6, n=8, L=120, seed=17)
    ctx, lib, P = w["ctx"], w["lib"], w["P"]
    b = ResidentBatch(ctx, w["mh"], w["handles"], w["node_off"], w["parent"], w["tip_row"])
    vals = b.eval(w["mh"], w["blen"])
    # a pending evaluation on one of the loci blocks a committing batch evaluation
    one = np.array([w["handles"][2]], dtype=np.uint64)
    a, e = int(w["node_off"][2]), int(w["node_off"][3])
    ctx.eval(w["mh"], one, np.array([0, e - a], dtype=np.int64), w["parent"][a:e], np.ascontiguousarray(w["blen"][a:e] * 1.5),
             w["tip_row"][a:e], 1.0, lib.EVAL_PENDING)
    with pytest.raises(PhyNetB200Error):
        b.eval(w["mh"], w["blen"])
    ctx.rollback(one)
    assert np.array_equal(b.eval(w["mh"], w["blen"]), vals)
    with pytest.raises(ValueError):
        ResidentBatch(ctx, w["mh"], np.concatenate([w["handles"][:2], w["handles"][:2]]),
                      np.array([0, 15, 30, 45, 60], dtype=np.int64), np.concatenate([w["parent"][:30]] * 2),
                      np.concatenate([w["tip_row"][:30]] * 2))        # a locus twice in a committing batch
    w["calcs"][4].close()                                            # a locus goes away: the batch is dead
    with pytest.raises(PhyNetB200Error):
        b.eval(w["mh"], w["blen"])
    # ... and the surviving loci keep the last values
    assert w["calcs"][1].log_likelihood(w["trees"][1], w["model"]) == vals[1]
    b.close()
