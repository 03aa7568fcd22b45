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
