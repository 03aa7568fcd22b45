"""GPU probe: where does the latency of one gene-tree move go?  pnb_eval (dirty path, one locus), pnb_msnc_eval (one gene
tree), pnb_eval_move (both), each through the Python binding; medians of 300 calls."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from phynetpy_b200 import infer as P, simulate as S, moves as MV, _lib
from phynetpy_b200._seq_likelihood import _model_handle
from phynetpy_b200._msnc_density import msnc_eval_packed
from phynetpy_b200.candidates import node_heights

rng = np.random.default_rng(45)
n_species, L, theta = 16, 500, 0.02
for n_retic in (0, 1, 2):
    net = S.random_species_network(n_species, n_retic, rng, height=0.005 * (n_species - 1))
    trees, species_of = S.simulate_gene_trees_msnc(net, 1, theta, rng, 8)
    aln = S.simulate_alignments(np.stack([t.parent for t in trees]), np.stack([t.blen for t in trees]), L, S.DEFAULT_PI, S.DEFAULT_RATES, rng)
    model = P.GTR(S.DEFAULT_PI, S.DEFAULT_RATES)
    calc = P.FelsensteinCalculator(S.to_alignment_dict(aln[0]))
    ctx = calc._context()
    mh = _model_handle(model, ctx)
    t = trees[0]
    h = node_heights(t)
    calc.log_likelihood(t, model)
    ids = net.species_ids_of(t, species_of)
    tip = t.tip_rows(calc._row_of)
    off = np.array([0, t.n_nodes], dtype=np.int64)
    nh = net.handle(ctx, None)
    lh = calc._locus_handle()
    harr = np.array([lh.value], dtype=np.uint64)
    props = []
    while len(props) < 300:
        mv = MV.slide_height_move(t, h, rng)
        if mv is not None:
            props.append(mv[0])
    res = {}
    for name in ("prune", "msnc", "move"):
        wall, abi, kern = [], [], []
        for k, p in enumerate(props):
            if k == 150:
                ctx.set_profiling(True)
            t0 = time.perf_counter()
            if name == "prune":
                ctx.eval(mh, harr, off, p.parent, p.blen, tip, 1.0, _lib.EVAL_PENDING)
            elif name == "msnc":
                msnc_eval_packed(net, off, p.parent, p.blen, ids, theta, ctx)
            else:
                ctx.eval_move(mh, lh, p.parent, p.blen, tip, 1.0, _lib.EVAL_PENDING, nh, ids, theta)
            wall.append(time.perf_counter() - t0)
            tm = ctx.last_timings()
            abi.append(tm["call_ms"])
            if k >= 150:
                kern.append(tm["prune_ms"])
            if name != "msnc":
                ctx.rollback(harr)
        ctx.set_profiling(False)
        res[name] = (np.median(wall) * 1e6, np.median(abi) * 1e3, np.median(kern) * 1e3)
    print("reticulations=%d " % n_retic + "  ".join("%s: wall %.1f us, C ABI %.1f us, kernel(events) %.1f us" % ((k,) + v) for k, v in res.items()), flush=True)
