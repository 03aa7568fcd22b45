"""K-m alone, for profiling: the candidate sets of 200 loci (16 species, two reticulations) scored
by pnb_msnc_eval a few times.  Used as
    ncu --set full --clock-control none --import-source on -k regex:pnb_msnc -s 2 -c 1 python tools/msnc_probe.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phynetpy_b200 import infer as P, simulate as S          # noqa: E402
from phynetpy_b200._msnc_density import msnc_eval_packed     # noqa: E402
from phynetpy_b200.candidates import candidate_arrays        # noqa: E402

rng = np.random.default_rng(44)
n_retic = int(os.environ.get("PROBE_RETIC", "2"))
net = S.random_species_network(16, n_retic, rng, height=0.075)
trees, species_of = S.simulate_gene_trees_msnc(net, 1, 0.02, rng, 200)
par, bl, ids, sizes = [], [], [], []
for t in trees:
    p, b = candidate_arrays(t)
    par.append(p.reshape(-1))
    bl.append(b.reshape(-1))
    ids.append(np.tile(net.species_ids_of(t, species_of), p.shape[0]))
    sizes.append(np.full(p.shape[0], p.shape[1], dtype=np.int64))
sizes = np.concatenate(sizes)
off = np.zeros(sizes.shape[0] + 1, dtype=np.int64)
np.cumsum(sizes, out=off[1:])
args = (off, np.ascontiguousarray(np.concatenate(par), dtype=np.int32), np.ascontiguousarray(np.concatenate(bl)),
        np.ascontiguousarray(np.concatenate(ids), dtype=np.int32))
if os.environ.get("PROBE_MODE") == "one":      # a single gene tree: the latency of one thread's DP
    t = trees[0]
    args = (np.array([0, t.n_nodes], dtype=np.int64), t.parent, t.blen, net.species_ids_of(t, species_of))
    off = args[0]
ctx = P.default_context()
for it in range(8):
    if it == 5:
        ctx.set_profiling(True)     # CUDA events around the kernel (slot "prune_ms" of last_timings)
    t0 = time.perf_counter()
    v = msnc_eval_packed(net, *args, 0.02, ctx)
    dt = time.perf_counter() - t0
    if it >= 5:
        print("kernel (events) %.1f us, call %.3f ms" % (ctx.last_timings()["prune_ms"] * 1e3, dt * 1e3))
ctx.set_profiling(False)
print("gene trees %d  finite %d  last call %.3f ms  stats %s" % (off.shape[0] - 1, int(np.isfinite(v).sum()), dt * 1e3, ctx.last_stats()))
