"""Worker of tests/test_gpu_multi.py (run under torch.distributed.run, one rank per GPU): the sharded
SeqLikelihoodEngine on real devices -- pruning kernel -> this rank's slice of the device vector -> in-place
ncclAllGather -> pinned host copy -- against the unsharded engine on rank 0, bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ["PHYNET_B200_DEVICE"] = str(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from phynetpy_b200 import infer as P, moves as MV, simulate as S
    from phynetpy_b200.candidates import node_heights
    from phynetpy_b200.distributed import ShardPlan

    rng = np.random.default_rng(31)
    M, n, L = 101, 12, 300                      # 101 loci: unequal shards, one padded slice
    par, bl = S.random_trees(M, n, rng)
    aln = S.simulate_alignments(par, bl, L, S.DEFAULT_PI, S.DEFAULT_RATES, rng)
    trees = S.flat_trees(par, bl)
    data = [S.to_alignment_dict(aln[i]) for i in range(M)]
    model = P.GTR(S.DEFAULT_PI, S.DEFAULT_RATES)

    def run(engine):
        r = np.random.default_rng(5)
        cur = list(trees)
        heights = [node_heights(t) for t in cur]
        total = engine.log_likelihood(cur, None, 0.0)
        trace = [total]
        for step in range(60):
            if step % 20 == 19:                 # an all-loci proposal: every length rescaled (resident batch path)
                f = float(np.exp(r.uniform(-0.05, 0.05)))
                prop = [t.with_lengths(t.blen * f) for t in cur]
                snap = engine.clone_caches()
                for i in range(M):
                    engine.invalidate_locus(i)
                new_total = engine.log_likelihood(prop, None, 0.0)
                if r.random() < 0.5:
                    cur, total = prop, new_total
                    heights = [h * f for h in heights]
                else:
                    engine.restore_caches(snap)
                trace.append(total)
                continue
            i = int(r.integers(M))
            mv = MV.slide_height_move(cur[i], heights[i], r)
            if mv is None:
                continue
            snap = engine.clone_caches()
            prop = list(cur)
            prop[i] = mv[0]
            engine.invalidate_locus(i)
            new_total = engine.log_likelihood(prop, None, 0.0)
            if r.random() < 0.5:
                cur, total = prop, new_total
                heights[i] = mv[1]
            else:
                engine.restore_caches(snap)
            trace.append(total)
        final = engine.log_likelihood(cur, None, 0.0)
        return list(engine._felsen), trace, final

    plan = ShardPlan.from_env(M)
    lo, hi = plan.local_range()
    results = []
    for resident in (True, False):          # all-loci proposals through a resident batch + fused peer stores / through pnb_eval
        sharded = P.SeqLikelihoodEngine([d if lo <= i < hi else None for i, d in enumerate(data)], None, model, shard=plan,
                                        resident_batches=resident)
        results.append(run(sharded))
    got = results[0]
    both_paths_agree = results[0] == results[1]
    # every rank holds the same values
    t = torch.tensor(got[0] + got[1] + [got[2]], dtype=torch.float64, device="cuda")
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    ok = same and both_paths_agree
    if rank == 0:
        single = P.SeqLikelihoodEngine(data, None, model)
        want = run(single)
        ok = ok and got[0] == want[0] and got[1] == want[1] and got[2] == want[2]
        print("SHARDED_ENGINE_%s ranks=%d same_on_all_ranks=%s equal_to_single_gpu=%s final=%.12f" %
              ("OK" if ok else "FAILED", world, same, got == want, got[2]), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
