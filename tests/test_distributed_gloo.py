"""CPU, world_size=2 over gloo: the N>1 host path -- contiguous locus shards, every rank
fills only its own range of the per-locus vector, ONE all-reduce(sum), host sums in locus
order.  The per-locus values come from the oracle here (no GPU); what is tested is the
sharding / combination logic that the GPU path uses unchanged (ShardPlan, balanced_ranges)."""
import json
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from oracle import felsenstein_oracle as O
    from phynetpy_b200.distributed import ShardPlan
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ec = json.load(open(os.path.join(ROOT, "tests", "golden", "engine_case.json")))
        model = O.gtr(ec["model"]["pi"], ec["model"]["rates"])
        M = len(ec["loci"])
        work = [len(l["alignment"][0][1]) * (len(l["tree"]["parent"]) // 2) for l in ec["loci"]]
        plan = ShardPlan.from_env(M, work)
        assert plan.world_size == world and plan.rank == rank
        lo, hi = plan.local_range()
        vec = np.zeros(M)
        for i in range(lo, hi):  # this rank's loci only
            l = ec["loci"][i]
            t = l["tree"]
            vec[i] = O.log_likelihood(O.OracleLocus(dict(l["alignment"])), O.OracleTree(t["parent"], t["blen"], t["label"]), model)
        full = plan.allreduce_sum(vec)
        total = 0.0
        for v in full:
            total += v
        q.put((rank, lo, hi, full.tolist(), total))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_sharded_sum_equals_single_rank_reference():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    ec = json.load(open(os.path.join(ROOT, "tests", "golden", "engine_case.json")))
    (r0, lo0, hi0, v0, t0), (r1, lo1, hi1, v1, t1) = res
    assert (lo0, hi1) == (0, len(ec["loci"])) and hi0 == lo1 and 0 < hi0 < hi1
    assert v0 == v1 and t0 == t1                      # every rank holds the same vector, bit for bit
    np.testing.assert_allclose(v0, ec["per_locus"], rtol=1e-12)
    assert t0 == pytest.approx(ec["total"], rel=1e-12)
    # identical to the single-process sum in locus order (adding zeros is exact)
    s = 0.0
    for v in v0:
        s += v
    assert s == t0


def _engine_worker(rank, world, port, q):
    """The sharded SeqLikelihoodEngine over gloo with the oracle-backed fake device of test_engine_host: the same
    chain (same seed) on every rank, each rank evaluates only its own loci, the per-locus vector is all-gathered in
    place.  Full evaluations, single-locus proposals, accepts and rejects."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from oracle import felsenstein_oracle as O
    from phynetpy_b200 import _engine as E, moves as MV, simulate as S
    from phynetpy_b200._seq_likelihood import GTR
    from phynetpy_b200.candidates import node_heights
    from phynetpy_b200.distributed import ShardPlan
    import test_engine_host as TH
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(23)
        M, n, L = 11, 6, 50                       # 11 loci over 2 ranks: unequal shards (5 + 6), one padded slice
        par, bl = S.random_trees(M, n, rng)
        aln = S.simulate_alignments(par, bl, L, S.DEFAULT_PI, S.DEFAULT_RATES, rng)
        trees = S.flat_trees(par, bl)
        data = [S.to_alignment_dict(aln[i]) for i in range(M)]
        om = O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES)
        E._model_handle = lambda model, ctx: None
        dev = TH.FakeDevice(data, om)
        plan = ShardPlan.from_env(M) if world > 1 else ShardPlan.single(M)
        lo, hi = plan.local_range()
        calcs = [TH._stub_calc(i, d) if lo <= i < hi else None for i, d in enumerate(data)]
        eng = E.SeqLikelihoodEngine(calcs, None, GTR(S.DEFAULT_PI, S.DEFAULT_RATES), context=dev, shard=plan)
        cur = list(trees)
        heights = [node_heights(t) for t in cur]
        total = eng.log_likelihood(cur, None, 0.0)
        trace = [total]
        evaluated = 0
        for step in range(40):
            i = int(rng.integers(M))
            mv = MV.slide_height_move(cur[i], heights[i], rng)
            if mv is None:
                continue
            snap = eng.clone_caches()
            prop = list(cur)
            prop[i] = mv[0]
            eng.invalidate_locus(i)
            n_eval = sum(c[0] == "eval" for c in dev.calls)
            new_total = eng.log_likelihood(prop, None, 0.0)
            evaluated += sum(c[0] == "eval" for c in dev.calls) - n_eval
            if rng.random() < 0.5:
                cur, total = prop, new_total
                heights[i] = mv[1]
            else:
                eng.restore_caches(snap)
            trace.append(total)
        final = eng.log_likelihood(cur, None, 0.0)
        q.put((rank, lo, hi, list(eng._felsen), trace, final, evaluated))
    finally:
        if world > 1:
            dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_sharded_engine_equals_the_single_process_engine():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_engine_worker, args=(r, 2, port, q)) for r in range(2)]
    procs.append(ctx.Process(target=_engine_worker, args=(0, 1, _free_port(), q)))
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    single = [r for r in res if (r[1], r[2]) == (0, 11)][0]
    two = sorted(r for r in res if (r[1], r[2]) != (0, 11))
    assert [(r[1], r[2]) for r in two] == [(0, 5), (5, 11)]      # unequal shards: one padded slice
    for r in two:
        assert r[3] == single[3]          # every rank holds every per-locus value, bit for bit the single-process ones
        assert r[4] == single[4] and r[5] == single[5]
    # each proposal was evaluated by the owner of its locus only
    assert two[0][6] + two[1][6] == single[6] and two[0][6] > 0 and two[1][6] > 0
