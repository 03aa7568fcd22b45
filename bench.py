#!/usr/bin/env python
"""bench.py -- site-node partial updates/s of the sequence-likelihood hot path.

    python bench.py --gpus 1 --steps K --warmup W              # cfg2: 50 taxa x 1e6 patterns, one GPU
    torchrun --nproc-per-node N bench.py --gpus N ...          # cfg5: 10,000 loci sharded over N GPUs
    python bench.py --impl reference ...                       # the reference algorithm on the host cores

One "step" = one full log-likelihood evaluation (every internal node recomputed,
P(t) of every branch, root reduction) of the workload.  Prints ONE JSON line.
See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_UPDATE = 96  # 2 child reads x 32 B + 1 write x 32 B
PI = (0.1, 0.2, 0.3, 0.4)
RATES = (1.0, 2.5, 0.7, 1.3, 3.1, 0.9)

WORKLOADS = {
    # name: (loci, taxa, sites or patterns, description)
    "cfg2": dict(M=1, n=50, P=1_000_000, desc="single 50-taxon gene tree, GTR, 1e6 unique site patterns"),
    "cfg3": dict(M=200, n=24, L=1000, desc="200 loci x 24 taxa x 1,000 sites, full logL eval"),
    "cfg4": dict(M=1000, n=16, L=500, desc="1,000 loci x 16 taxa x 500 sites"),
    "cfg5": dict(M=10_000, n=32, L=2000, desc="10,000 loci x 32 taxa x 2,000 sites sharded over N GPUs, NCCL logL allreduce"),
}


# --------------------------------------------------------------------------- helpers
def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_locus_data(M, n, L, seed, lo=0, hi=None, chunk=250):
    """Seeded synthetic loci [lo, hi): trees + (n, L) ASCII alignments (per-locus seeds so any
    shard of the same workload sees identical data regardless of the number of ranks)."""
    from phynetpy_b200 import simulate as S
    hi = M if hi is None else hi
    trees, alns = [], []
    for c0 in range(lo, hi, chunk):
        c1 = min(hi, c0 + chunk)
        rng = np.random.default_rng([seed, c0])
        par, bl = S.random_trees(c1 - c0, n, rng)
        aln = S.simulate_alignments(par, bl, L, PI, RATES, rng)
        trees.extend(S.flat_trees(par, bl))
        alns.extend(aln[i] for i in range(c1 - c0))
    return trees, alns


def build_cfg2(P_target, n, seed, P):
    """50 taxa, exactly P_target unique patterns: simulate, compress on the GPU, keep the first
    P_target patterns with their weights (cfg 2)."""
    from phynetpy_b200 import simulate as S
    from phynetpy_b200._seq_likelihood import compress_alignment, _CODE_OF_BYTE
    rng = np.random.default_rng(seed)
    par, bl = S.random_trees(1, n, rng)
    ft = S.flat_trees(par, bl)[0]
    L = int(P_target * 1.06) + 64
    blocks = []
    for c0 in range(0, L, 131072):  # bounded host memory while simulating
        blocks.append(S.simulate_alignments(par, bl, min(131072, L - c0), PI, RATES, rng)[0])
    mat = np.concatenate(blocks, axis=1)
    t0 = time.perf_counter()
    codes, w, first, _ = compress_alignment(mat, _CODE_OF_BYTE, use_device=True)
    t_comp = time.perf_counter() - t0
    if codes.shape[1] < P_target:
        raise RuntimeError("only %d unique patterns from %d sites" % (codes.shape[1], L))
    codes = np.ascontiguousarray(codes[:, :P_target])
    w = w[:P_target].copy()
    fc = P.FelsensteinCalculator.from_codes(S.tip_labels(n), codes, w)
    return fc, ft, {"sites_simulated": L, "device_compress_s": round(t_comp, 4)}


# --------------------------------------------------------------------------- CPU legs
def oracle_time_single(codes, weights, ft, n_patterns, repeats):
    """The reference algorithm (oracle port, NumPy) on the first n_patterns patterns."""
    from oracle import felsenstein_oracle as O
    om = O.gtr(PI, RATES)
    tree = O.OracleTree(ft.parent.tolist(), [None if math.isnan(b) else float(b) for b in ft.blen], ft.label)
    tip_row = ft.tip_rows({"t%d" % i: i for i in range(codes.shape[0])})
    locus = O.OracleCodesLocus(np.ascontiguousarray(codes[:, :n_patterns]), weights[:n_patterns])  # construction untimed
    best = float("inf")
    val = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        val = locus.log_likelihood(tip_row, tree, om)
        best = min(best, time.perf_counter() - t0)
    n_int = int((ft.n_children() > 0).sum())
    return best, n_int * n_patterns, val


def _oracle_worker(args):
    from oracle import felsenstein_oracle as O
    alns, trees, reps = args
    om = O.gtr(PI, RATES)
    loci = [O.OracleLocus({"t%d" % i: a[i].tobytes().decode() for i in range(a.shape[0])}) for a in alns]
    otrees = [O.OracleTree(t.parent.tolist(), [None if math.isnan(b) else float(b) for b in t.blen], t.label) for t in trees]
    t0 = time.perf_counter()
    for _ in range(reps):
        vals = [O.log_likelihood(l, t, om) for l, t in zip(loci, otrees)]
    dt = (time.perf_counter() - t0) / reps
    ups = sum(l.n_patterns * (len(t.parent) - sum(1 for c in t.children if not c)) for l, t in zip(loci, otrees))
    return dt, ups, vals


def oracle_time_loci(trees, alns, procs, reps=10):
    """Reference loop over loci (src/_mcmc_seq.py:746-750), loci split over host processes."""
    import multiprocessing as mp
    procs = max(1, min(procs, len(trees)))
    parts = [(alns[i::procs], trees[i::procs], reps) for i in range(procs)]
    if procs == 1:
        res = [_oracle_worker(parts[0])]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_oracle_worker, parts)
    wall = max(r[0] for r in res)  # evaluation only; construction is excluded as in BASELINE.md
    ups = sum(r[1] for r in res)
    return wall, ups


# --------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (smoke runs only; reported)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-allreduce", action="store_true", help="diagnostic only: time the sharded step without the collective")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = args.workload
    if wl == "auto":
        wl = "cfg2" if max(args.gpus, world) == 1 else "cfg5"
    spec = dict(WORKLOADS[wl])
    if args.scale != 1.0:
        for k in ("P", "M"):
            if k in spec and spec[k] > 1:
                spec[k] = max(1, int(spec[k] * args.scale))

    if args.impl == "reference":
        return run_reference(args, wl, spec, rank)

    import torch
    import torch.distributed as dist
    from phynetpy_b200 import infer as P
    from phynetpy_b200 import _lib
    from phynetpy_b200._seq_likelihood import _model_handle
    from phynetpy_b200.device import pack_trees
    from phynetpy_b200.distributed import ShardPlan, balanced_ranges

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    os.environ["PHYNET_B200_DEVICE"] = str(local_rank)
    ctx = P.default_context(local_rank)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_profiling(True)
    model = P.GTR(PI, RATES)
    mh = _model_handle(model, ctx)
    setup = {}

    # ---- build the resident workload -------------------------------------------------
    t_setup0 = time.perf_counter()
    if wl == "cfg2":
        if world > 1:
            raise SystemExit("cfg2 is a single-locus, single-GPU workload (replicas only); use cfg5 for N>1")
        fc, ft, setup = build_cfg2(spec["P"], spec["n"], 2, P)
        calcs, trees = [fc], [ft]
        M = 1
        lo, hi = 0, 1
        plan = ShardPlan.single(1)
        cpu_sample = (fc._codes, fc._int_weights, ft)
    else:
        M = spec["M"]
        ranges = balanced_ranges(np.ones(M), world)
        plan = ShardPlan(M, rank, world, ranges)
        lo, hi = plan.local_range()
        trees, alns = make_locus_data(M, spec["n"], spec["L"], {"cfg3": 3, "cfg4": 4, "cfg5": 5}[wl], lo, hi)
        taxa = ["t%d" % i for i in range(spec["n"])]
        calcs = [P.FelsensteinCalculator.from_matrix(taxa, a, device_compress=False) for a in alns]
        cpu_sample = (trees[:256], alns[:256])
        del alns
    for c in calcs:
        c._locus_handle()
    handles = np.array([c._locus.value for c in calcs], dtype=np.uint64)
    node_off, parent, blen, tip_row = pack_trees(trees, [c._row_of for c in calcs])
    n_local = len(calcs)
    patterns_local = int(sum(c.n_patterns for c in calcs))
    setup["setup_s"] = round(time.perf_counter() - t_setup0, 2)

    out_vec = torch.zeros(M, dtype=torch.float64, device="cuda")      # all-reduced per-locus logL
    local_vec = torch.zeros(M, dtype=torch.float64, device="cuda") if world > 1 else out_vec
    out_local_ptr = local_vec.data_ptr() + 8 * lo                      # the kernel writes this rank's slice
    flags = _lib.EVAL_FULL

    # every step scores the same topologies with NEW branch lengths (a height-rescaling proposal):
    # nothing of a previous step's result can be reused, and the host recompiles lengths each time
    blens = [np.ascontiguousarray(blen * (1.0 + 0.002 * k)) for k in range(max(min(args.steps, 8), 1))]
    step_no = [0]

    def next_blen():
        b = blens[step_no[0] % len(blens)]
        step_no[0] += 1
        return b

    def step_device():
        """inputs resident in HBM; logL stays on the device; all-reduce when sharded"""
        with torch.cuda.stream(stream):
            ctx.eval(mh, handles, node_off, parent, next_blen(), tip_row, 1.0, flags, None, out_device_ptr=out_local_ptr)
            if world > 1 and not args.skip_allreduce:
                out_vec.copy_(local_vec)   # zeros outside this rank's range: the sum is exact
                dist.all_reduce(out_vec)

    def step_e2e():
        """the call a user makes: host trees in, host log-likelihoods out"""
        vals = ctx.eval(mh, handles, node_off, parent, next_blen(), tip_row, 1.0, flags, None)
        if world > 1:
            v = torch.zeros(M, dtype=torch.float64)
            v[lo:hi] = torch.from_numpy(vals)
            v = v.cuda()
            dist.all_reduce(v)
            return v.cpu().numpy()
        return vals

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    stats = ctx.last_stats()
    updates_local = stats["updates"]
    launches_per_step = stats["launches"]

    # ---- timed: device-resident ---------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ctx.set_profiling(False)
    barrier()
    step_no[0] = 0
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    logL_device = out_vec.cpu().numpy().copy()

    # ---- per-kernel durations: same steps again with CUDA events around each kernel launch ----
    ctx.set_profiling(True)
    prune_ms, pm_ms, comp_ms, call_ms = [], [], [], []
    for _ in range(args.steps):
        step_device()
        tm = ctx.last_timings()  # waits for this step's events
        prune_ms.append(tm["prune_ms"]); pm_ms.append(tm["pmatrix_ms"]); comp_ms.append(tm["compile_ms"])
        call_ms.append(tm["call_ms"])
    barrier()

    # ---- score-only mode (PNB_EVAL_NOSTORE): same evaluation, partials never leave the SM ----
    score_ms = None
    if wl == "cfg2" or world == 1:
        for c in calcs:
            c.invalidate()
        sflags = _lib.EVAL_NOSTORE
        def step_score():
            with torch.cuda.stream(stream):
                ctx.eval(mh, handles, node_off, parent, next_blen(), tip_row, 1.0, sflags, None, out_device_ptr=out_local_ptr)
        for _ in range(3):
            step_score()
        barrier()
        sk = []
        step_no[0] = 0
        for _ in range(args.steps):
            step_score()
            sk.append(ctx.last_timings()["prune_ms"])
        barrier()
        score_ms = float(np.mean(sk))
        score_updates = ctx.last_stats()["updates"]
        score_logL = out_vec.cpu().numpy().copy()
    ctx.set_profiling(False)

    # ---- cfg4: MCMC chain steps with dirty-path recompute (one locus per proposal) -------------
    chain = None
    if wl == "cfg4" and world == 1:
        chain = chain_steps(P, ctx, calcs, trees, model, n_steps=min(max(200, 20 * args.steps), 2000))
        chain["coupled_move_candidates"] = candidate_batch(P, ctx, calcs, trees, model, n_loci=min(200, len(calcs)))
        chain["msnc_density"] = msnc_leg(P, ctx, n_loci=len(calcs), n_species=spec["n"], cpu=not args.no_cpu_baseline)
        chain["coupled_move"] = coupled_move_leg(P, ctx, model, n_loci=min(200, len(calcs)), n_species=spec["n"], L=spec["L"])
        chain["joint_likelihood_chain"] = joint_chain(P, ctx, model, n_loci=len(calcs), n_species=spec["n"], L=spec["L"],
                                                      n_steps=1000)

    # ---- timed: end to end through the C ABI with host buffers ---------------------------
    for _ in range(2):
        step_e2e()
    barrier()
    step_no[0] = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals_e2e = step_e2e()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    st_e2e = ctx.last_stats()
    # the same call one level up: the reference-shaped Python facade with a Network-like tree OBJECT
    # (root()/get_children()/get_edge().get_length()), marshalled per call as the reference's callers do
    facade_ms = None
    if wl == "cfg2":
        gt = P.GeneTree.from_flat(trees[0])
        calcs[0].log_likelihood(gt, model, full=True)
        t0 = time.perf_counter()
        nrep = min(args.steps, 10)
        for _ in range(nrep):
            facade_val = calcs[0].log_likelihood(gt, model, full=True)
        facade_ms = (time.perf_counter() - t0) * 1e3 / nrep
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # ---- reduce over ranks ----------------------------------------------------------------
    t = torch.tensor([elapsed_ms, e2e_ms], dtype=torch.float64, device="cuda")
    u = torch.tensor([float(updates_local), float(patterns_local)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    elapsed_ms, e2e_ms = float(t[0]), float(t[1])
    updates_total = float(u[0])

    if rank == 0:
        ms_per_step = elapsed_ms / args.steps
        value = updates_total * args.steps / (elapsed_ms * 1e-3)
        e2e_value = updates_total * args.steps / (e2e_ms * 1e-3)
        peak, peak_src = peak_hbm()
        k_ms = float(np.mean(prune_ms))
        achieved = updates_local * ALGO_BYTES_PER_UPDATE / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_%s.json" % wl)
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        total_logL = float(np.sum(logL_device)) if world > 1 or wl != "cfg2" else float(logL_device[0])
        e2e_total = float(np.sum(vals_e2e))
        result = {
            "metric": "site-node partial updates/sec",
            "value": value,
            "unit": "updates/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "strong" if wl != "cfg2" else "weak",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": "%s: %s" % (wl, spec["desc"]),
                "loci": M, "taxa": spec["n"], "site_patterns_total": int(u[1]),
                "updates_per_step": int(updates_total),
                "model": "GTR pi=(.1,.2,.3,.4) rates=(1,2.5,.7,1.3,3.1,.9)",
                "step": "full logL evaluation of a fresh proposal (same topologies, new branch lengths every step): "
                        "P(t) of every branch + every internal node recomputed + root reduction"
                        + (" + NCCL all-reduce of the per-locus vector" if world > 1 else ""),
                "parallelism": "loci sharded over %d GPU(s), one process per GPU" % world,
                "l2": "working set per step (%.2f GB of partials written) exceeds the 126 MB L2; no flush needed"
                      % (updates_local * 32 / 1e9),
                "scale": args.scale, "diagnostic_skip_allreduce": bool(args.skip_allreduce),
                **setup,
            },
            "e2e": {
                "value": e2e_value, "unit": "updates/s", "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": int(st_e2e["h2d_bytes"]), "d2h_bytes_per_step": int(st_e2e["d2h_bytes"]),
                "what": "pnb_eval through the C ABI: host tree arrays in (compiled to the op program on the host, "
                        "copied H2D), kernels, per-locus logL copied D2H; alignment resident since construction",
                "python_facade_ms_per_step": facade_ms,
                "python_facade_what": None if facade_ms is None else
                    "FelsensteinCalculator.log_likelihood(tree object, model, full=True): per-call marshalling of a "
                    "Network-like object included",
            },
            "gpu_launches": int(launches_per_step * args.steps),
            "kernels_per_step": {"pnb_pmatrix_kernel": 1, "pnb_prune_kernel": 1},
            "roofline": {
                "bound": "hbm", "kernel": "pnb_prune_kernel",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None,
                "traffic": traffic,
                "peak_source": peak_src,
                "algorithmic_bytes_per_update": ALGO_BYTES_PER_UPDATE,
                "kernel_ms": k_ms, "pmatrix_kernel_ms": float(np.mean(pm_ms)),
                "host_compile_ms": float(np.mean(comp_ms)), "host_call_ms": float(np.mean(call_ms)),
                "note": "achieved uses the 96 B/update streaming model; the kernel keeps a node's freshly computed "
                        "child in registers, so its real DRAM traffic (`traffic`) is lower than algorithmic",
            },
            "score_only": None if score_ms is None else {
                "what": "PNB_EVAL_NOSTORE: every node recomputed, no partial written to HBM (candidate scoring)",
                "kernel_ms": score_ms, "updates_per_s_kernel": score_updates / (score_ms * 1e-3),
                "logL_equal_to_store_mode": bool(np.array_equal(score_logL, logL_device)),
            },
            "chain": chain,
            "clocks": clocks,
            "logL": total_logL,
            "logL_e2e": e2e_total,
        }
        if not args.no_cpu_baseline:
            result["cpu_baseline"] = cpu_baseline(wl, spec, cpu_sample)
            chk = result["cpu_baseline"].pop("_check", None)
            if chk is not None:
                result["parity"] = chk(calcs, trees, model, P)
        print(json.dumps(result))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def chain_steps(P, ctx, calcs, trees, model, n_steps, seed=4):
    """BASELINE cfg4: a stream of single-locus proposals through the engine seam the sampler uses
    (invalidate_locus -> log_likelihood -> accept | restore_caches).  Each proposal moves one
    internal node's height (its two child branches and its parent branch change, as
    op_gene_node_height does, reference src/_mcmc_seq.py:1020-1039), so the device recomputes
    that node and its ancestors only.  Ends with a full recomputation of every locus, which must
    reproduce the incrementally maintained total bit for bit."""
    rng = np.random.default_rng(seed)
    eng = P.SeqLikelihoodEngine(calcs, None, model, context=ctx)
    cur = list(trees)
    total = eng.log_likelihood(cur, None, 0.0)
    eng._commit_pending()
    M = len(cur)
    nodes = cur[0].n_nodes
    n_tips = (nodes + 1) // 2
    upd, recomputed, accepted = 0, 0, 0
    lat, c_ms = [], []
    t0 = time.perf_counter()
    for _ in range(n_steps):
        i = int(rng.integers(0, M))
        ft = cur[i]
        v = int(rng.integers(n_tips, nodes - 1))           # internal, non-root
        kids = np.nonzero(ft.parent == v)[0]
        room_dn = float(ft.blen[kids].min())
        room_up = float(ft.blen[v])
        delta = float(rng.uniform(-0.5 * room_dn, 0.5 * room_up))
        bl = ft.blen.copy()
        bl[kids] += delta
        bl[v] -= delta
        prop = list(cur)
        prop[i] = ft.with_lengths(bl)
        snap = eng.clone_caches()
        eng.invalidate_locus(i)
        s0 = time.perf_counter()
        new_total = eng.log_likelihood(prop, None, 0.0)
        lat.append(time.perf_counter() - s0)
        c_ms.append(ctx.last_timings()["call_ms"])
        upd += eng.last_stats["updates"]
        recomputed += eng.last_stats["nodes_recomputed"]
        if math.log(rng.random()) < (new_total - total):    # Metropolis on the likelihood alone
            cur, total = prop, new_total
            accepted += 1
        else:
            eng.restore_caches(snap)
    wall = time.perf_counter() - t0
    eng._commit_pending()
    incremental = list(eng._felsen)
    # full recomputation of the final state
    for i in range(M):
        eng.invalidate_locus(i)
    eng.invalidate_device_cache()
    full = eng.log_likelihood(cur, None, 0.0)
    eng._commit_pending()
    return {
        "what": "single-locus node-height proposals through SeqLikelihoodEngine (dirty-path recompute, "
                "commit on accept / rollback on reject), Python host loop included",
        "steps": n_steps, "steps_per_s": n_steps / wall, "accept_rate": accepted / n_steps,
        "mean_nodes_recomputed": recomputed / n_steps, "mean_updates_per_step": upd / n_steps,
        "median_eval_latency_us": float(np.median(lat) * 1e6),
        "median_c_abi_call_us": float(np.median(c_ms) * 1e3),
        "full_recompute_total": full,
        "per_locus_incremental_equals_full_recompute_bitwise": bool(incremental == list(eng._felsen)),
    }


def joint_chain(P, ctx, model, n_loci, n_species, L, n_steps, seed=45):
    """MCMC_SEQ's whole likelihood, sum_i [log P(S_i|g_i) + log P(g_i|Psi)] (reference src/_mcmc_seq.py:745-761),
    through the engine seam with BOTH factors on the device: single-locus node-height proposals (dirty-path
    pruning + one MSNC density) and, every 10th step, a theta proposal (invalidate_network: the MSNC density of
    every locus in one call).  Data: a 1-reticulation species network, gene trees from the network coalescent on
    it, GTR alignments on those trees.  Ends with a full recomputation that must reproduce both cached factors."""
    from phynetpy_b200 import simulate as S
    rng = np.random.default_rng(seed)
    theta = 0.02
    net = S.random_species_network(n_species, 1, rng, height=0.005 * (n_species - 1))
    trees, species_of = S.simulate_gene_trees_msnc(net, 1, theta, rng, n_loci)
    par = np.stack([t.parent for t in trees])
    bl = np.stack([t.blen for t in trees])
    aln = S.simulate_alignments(par, bl, L, S.DEFAULT_PI, S.DEFAULT_RATES, rng)
    loci = [S.to_alignment_dict(aln[i]) for i in range(n_loci)]
    eng = P.SeqLikelihoodEngine(loci, species_of, model, context=ctx)
    cur = list(trees)
    total = eng.log_likelihood(cur, net, theta)
    eng._commit_pending()
    from phynetpy_b200 import moves as MV
    from phynetpy_b200.candidates import node_heights
    heights = [node_heights(t) for t in cur]
    lat_locus, lat_theta, accepted, impossible, void = [], [], 0, 0, 0
    t0 = time.perf_counter()
    for step in range(n_steps):
        snap = eng.clone_caches()
        prop_h, loghr = None, 0.0
        if step % 10 == 9:
            new_theta = theta * math.exp(float(rng.uniform(-0.2, 0.2)))
            loghr = math.log(new_theta / theta)            # scaler move (op_change_theta, src/_mcmc_seq.py:888-915)
            eng.invalidate_network()
            prop = cur
            s0 = time.perf_counter()
            new_total = eng.log_likelihood(prop, net, new_theta)
            lat_theta.append(time.perf_counter() - s0)
        else:
            new_theta = theta
            i = int(rng.integers(0, n_loci))
            # the reference's gene-tree operators on flat arrays (phynetpy_b200/moves.py): node-height slide
            # (op_gene_node_height) or height-preserving NNI (op_gene_tree_nni)
            if step % 3 == 2:
                out = MV.nni_move(cur[i], heights[i], rng)
                move = None if out is None else (out[0], out[1], 0.0)
            else:
                move = MV.slide_height_move(cur[i], heights[i], rng)
            if move is None:
                void += 1
                continue
            prop = list(cur)
            prop[i] = move[0]
            prop_h = (i, move[1])
            loghr = move[2]
            eng.invalidate_locus(i)
            s0 = time.perf_counter()
            new_total = eng.log_likelihood(prop, net, theta)
            lat_locus.append(time.perf_counter() - s0)
        impossible += new_total == -math.inf
        if math.log(rng.random()) < (new_total - total) + loghr:
            cur, total, theta = prop, new_total, new_theta
            if prop_h is not None:
                heights[prop_h[0]] = prop_h[1]
            accepted += 1
        else:
            eng.restore_caches(snap)
    wall = time.perf_counter() - t0
    eng._commit_pending()
    inc_f, inc_m = list(eng._felsen), list(eng._msnc)
    for i in range(n_loci):
        eng.invalidate_locus(i)
    eng.invalidate_device_cache()
    full = eng.log_likelihood(cur, net, theta)
    eng._commit_pending()
    return {
        "what": "both factors of the MCMC_SEQ likelihood on the device through SeqLikelihoodEngine.log_likelihood("
                "gene_trees, species_net, theta); Python host loop included",
        "loci": n_loci, "species": n_species, "reticulations": 1, "sites": L,
        "steps": n_steps, "logL_evaluations_per_s": n_steps / wall, "accept_rate": accepted / n_steps,
        "proposals_without_embedding": int(impossible), "void_proposals": int(void),
        "moves": "node-height slide / height-preserving NNI of one locus (moves.py), theta scaling every 10th step",
        "median_latency_us_locus_move": float(np.median(lat_locus) * 1e6),
        "median_latency_us_theta_move_all_loci_msnc": float(np.median(lat_theta) * 1e6),
        "final_total": full,
        "incremental_equals_full_recompute_bitwise": bool(inc_f == list(eng._felsen) and inc_m == list(eng._msnc)
                                                          and full == total),
    }


def coupled_move_leg(P, ctx, model, n_loci, n_species, L, seed=46):
    """One guided joint gene-tree re-proposal of a coupled add-reticulation move (reference
    _coupled_gene_tree_reproposal, src/_mcmc_seq.py:1995-2070): every locus' candidate set scored under the target
    network (Felsenstein + MSNC), one member sampled, the candidate sets of the chosen trees scored under the
    reverse network -- four device calls in phynetpy_b200/coupled.py; host enumeration included."""
    from phynetpy_b200 import simulate as S
    from phynetpy_b200.candidates import candidate_table, node_heights
    from phynetpy_b200.coupled import coupled_gene_tree_reproposal
    rng = np.random.default_rng(seed)
    theta = 0.02
    sp_tree = S.random_species_network(n_species, 0, rng, height=0.005 * (n_species - 1))
    # the same tree with one reticulation = the network the move proposes
    rng2 = np.random.default_rng(seed)
    sp_net = S.random_species_network(n_species, 1, rng2, height=0.005 * (n_species - 1))
    trees, species_of = S.simulate_gene_trees_msnc(sp_tree, 1, theta, rng, n_loci)
    aln = S.simulate_alignments(np.stack([t.parent for t in trees]), np.stack([t.blen for t in trees]), L,
                                S.DEFAULT_PI, S.DEFAULT_RATES, rng)
    eng = P.SeqLikelihoodEngine([S.to_alignment_dict(aln[i]) for i in range(n_loci)], species_of, model, context=ctx)
    eng.log_likelihood(trees, sp_tree, theta)
    eng._commit_pending()
    heights = [node_heights(t) for t in trees]
    res, times = None, []
    for rep in range(4):
        t0 = time.perf_counter()
        res = coupled_gene_tree_reproposal(eng, trees, heights, sp_net, sp_tree, theta, np.random.default_rng(rep))
        times.append(time.perf_counter() - t0)
    n_members = sum(candidate_table(t, h)[0].shape[0] for t, h in zip(trees, heights))
    return {
        "what": "coupled_gene_tree_reproposal: forward + reverse candidate sets of all loci, 4 device calls",
        "loci": n_loci, "forward_candidate_trees": int(n_members),
        "prunings_and_densities_the_reference_would_run": int(4 * n_members),
        "seconds_per_move": float(np.median(times[1:])), "first_call_seconds": float(times[0]),
        "declined": res is None,
        "log_qf": None if res is None else res[2], "log_qr": None if res is None else res[3],
    }


def candidate_batch(P, ctx, calcs, trees, model, n_loci):
    """The coupled moves' inner double loop (reference src/_mcmc_seq.py:2028 x :1935-1957): for every locus
    the candidate set (tree + height-preserving NNI neighbours) x (0.25, 0.5, 1, 2, 4), every member scored
    by a full pruning -- here all loci x all candidates in ONE PNB_EVAL_NOSTORE call."""
    from phynetpy_b200.candidates import candidate_arrays
    eng = P.SeqLikelihoodEngine(calcs[:n_loci], None, model, context=ctx)
    base = eng.log_likelihood(trees[:n_loci], None, 0.0)
    eng._commit_pending()
    t0 = time.perf_counter()
    sets = [(i, (trees[i],) + candidate_arrays(trees[i])) for i in range(n_loci)]
    t_enum = time.perf_counter() - t0
    n_trees = sum(c[1].shape[0] for _, c in sets)
    for _ in range(2):                                 # warm-up: templates, both pinned staging buffers
        eng.score_candidate_sets(sets)
    t0 = time.perf_counter()
    res = eng.score_candidate_sets(sets)
    wall = time.perf_counter() - t0
    st = dict(eng.last_stats)
    tm = ctx.last_timings()
    ctx.set_profiling(True)
    eng.score_candidate_sets(sets)
    tm_prof = ctx.last_timings()
    ctx.set_profiling(False)
    # the identity candidate of every locus must reproduce the committed tree's value
    ident_ok = all(any(v == eng._felsen[i] for v in vals) for (i, _), vals in zip(sets, res))
    return {
        "what": "all candidates of %d loci scored in one device call (NOSTORE), Python packing included" % n_loci,
        "candidate_trees": n_trees, "trees_per_locus": n_trees / n_loci,
        "seconds": wall, "trees_per_s": n_trees / wall, "updates_per_s": st["updates"] / wall,
        "c_abi_call_ms": tm["call_ms"], "host_compile_ms": tm["compile_ms"],
        "kernel_span_ms": tm_prof["prune_ms"], "tiles": st["tiles"], "launches": st["launches"],
        "enumeration_seconds_python": t_enum,
        "identity_candidate_reproduces_committed_value": bool(ident_ok), "base_total": base,
    }


def msnc_leg(P, ctx, n_loci, n_species, cpu=True, seed=44):
    """The other factor of the MCMC_SEQ likelihood, log P(g_i | Psi) (kernel K-m):
    all loci of a full evaluation in one pnb_msnc_eval call, and the candidate sets of the coupled moves in
    one call, on a species tree and on a network with two reticulations.  Host arrays in, host values out."""
    from phynetpy_b200 import simulate as S
    from phynetpy_b200._msnc_density import msnc_eval_packed
    from phynetpy_b200.candidates import candidate_arrays
    from oracle import msnc_oracle as MO   # checker and CPU baseline only
    rng = np.random.default_rng(seed)
    theta = 0.02
    out = {}
    for name, n_retic in (("species_tree", 0), ("network_2_reticulations", 2)):
        net = S.random_species_network(n_species, n_retic, rng, height=0.005 * (n_species - 1))
        trees, species_of = S.simulate_gene_trees_msnc(net, 1, theta, rng, n_loci)

        def pack(sets):
            par = np.ascontiguousarray(np.concatenate([p.reshape(-1) for p, _, _ in sets]), dtype=np.int32)
            bl = np.ascontiguousarray(np.concatenate([b.reshape(-1) for _, b, _ in sets]), dtype=np.float64)
            ids = np.ascontiguousarray(np.concatenate([np.tile(i, p.shape[0]) for p, _, i in sets]), dtype=np.int32)
            sizes = np.concatenate([np.full(p.shape[0], p.shape[1], dtype=np.int64) for p, _, _ in sets])
            off = np.zeros(sizes.shape[0] + 1, dtype=np.int64)
            np.cumsum(sizes, out=off[1:])
            return off, par, bl, ids

        ids_of = [net.species_ids_of(t, species_of) for t in trees]
        full = pack([(t.parent[None, :], t.blen[None, :], i) for t, i in zip(trees, ids_of)])
        n_cand_loci = min(200, n_loci)
        cand = pack([candidate_arrays(trees[k]) + (ids_of[k],) for k in range(n_cand_loci)])
        res = {"reticulations": n_retic, "program": net.info(ctx)}
        one = pack([(trees[0].parent[None, :], trees[0].blen[None, :], ids_of[0])])
        for tag, arrs in (("one_locus", one), ("all_loci", full), ("candidate_sets", cand)):
            n_trees = arrs[0].shape[0] - 1
            for _ in range(3):
                vals = msnc_eval_packed(net, *arrs, theta, ctx)
            reps = 200 if tag == "one_locus" else 20
            t0 = time.perf_counter()
            for _ in range(reps):
                vals = msnc_eval_packed(net, *arrs, theta, ctx)
            wall = (time.perf_counter() - t0) / reps
            ctx.set_profiling(True)
            msnc_eval_packed(net, *arrs, theta, ctx)
            tm = ctx.last_timings()
            ctx.set_profiling(False)
            res[tag] = {"gene_trees": int(n_trees), "ms_per_call": wall * 1e3, "trees_per_s": n_trees / wall,
                        "kernel_ms": tm["prune_ms"], "host_prepare_ms": tm["compile_ms"],
                        "finite": int(np.isfinite(vals).sum())}
            if tag == "all_loci":
                netd = net.to_dict()
                k = min(100, n_loci)
                t0 = time.perf_counter()
                want = [MO.msnc_log_density(netd, t.parent.tolist(), t.blen.tolist(), t.label, species_of, theta)
                        for t in trees[:k]] if cpu else []
                dt = time.perf_counter() - t0
                if cpu:
                    err = max(abs(g - w) / max(1.0, abs(w)) for g, w in zip(vals[:k], want))
                    res["parity_max_rel_err_vs_oracle"] = err
                    res["cpu_baseline"] = {"value": k / dt, "unit": "gene trees/s", "cores": 1, "kind": "port",
                                           "sample": "oracle/msnc_oracle.py (Python restatement of the reference DP) on %d of the "
                                                     "gene trees" % k}
        out[name] = res
    out["what"] = ("pnb_msnc_eval through the C ABI: packed host tree arrays in -> host preparation (heights, sorted "
                   "coalescence events, lineage sets) -> one kernel, one thread per gene tree -> values out")
    return out


def cpu_baseline(wl, spec, sample):
    cores = os.cpu_count() or 1
    if wl == "cfg2":
        codes, weights, ft = sample
        n_pat = min(200_000, codes.shape[1])
        dt, ups, val = oracle_time_single(codes, weights, ft, n_pat, repeats=3)

        def check(calcs, trees, model, P):
            sub = P.FelsensteinCalculator.from_codes(calcs[0].taxa, codes[:, :n_pat], weights[:n_pat])
            got = sub.log_likelihood(trees[0], model)
            sub.close()
            return {"oracle_logL_sample": val, "gpu_logL_sample": got, "rel_err": abs(got - val) / abs(val),
                    "sample": "first %d patterns" % n_pat}
        return {"value": ups / dt, "unit": "updates/s", "cores": cores, "kind": "port",
                "sample": "oracle (NumPy restatement of the reference's _postorder_partials) on the first %d of the "
                          "patterns, best of 3 evaluations; NumPy/OpenBLAS default threading" % n_pat,
                "seconds_per_eval_sample": dt, "_check": check}
    trees, alns = sample
    procs = min(cores, 32)
    dt, ups = oracle_time_loci(trees, alns, procs)
    return {"value": ups / dt, "unit": "updates/s", "cores": procs, "kind": "port",
            "sample": "oracle (NumPy restatement of the reference loop src/_mcmc_seq.py:746-750) on the first %d loci, "
                      "loci split over %d processes, mean of 10 evaluations each (construction excluded)" % (len(trees), procs),
            "seconds_sample": dt}


def run_reference(args, wl, spec, rank):
    """The reference's CPU algorithm for the path (oracle port: the reference is pure NumPy/Python
    and cannot travel to the GPU box) on this box's host cores, bounded sample per step."""
    if rank != 0:
        return
    from phynetpy_b200 import simulate as S
    from oracle import felsenstein_oracle as O
    cores = os.cpu_count() or 1
    steps, warm = args.steps, max(args.warmup, 1)
    if wl == "cfg2":
        n, n_pat = spec["n"], 100_000
        rng = np.random.default_rng(2)
        par, bl = S.random_trees(1, n, rng)
        ft = S.flat_trees(par, bl)[0]
        mat = S.simulate_alignments(par, bl, int(n_pat * 1.05), PI, RATES, rng)[0]
        first, w, _ = O.compress_fast(mat)
        lut = np.array([O.tip_code_for_char(chr(b)) for b in range(256)], dtype=np.uint8)
        codes = lut[mat[:, first[:n_pat]]]
        w = w[:n_pat]
        n_pat = codes.shape[1]
        steps = min(steps, 10)
        for _ in range(warm):
            oracle_time_single(codes, w, ft, n_pat, 1)
        t0 = time.perf_counter()
        ups = 0
        for _ in range(steps):
            _, u, _ = oracle_time_single(codes, w, ft, n_pat, 1)
            ups += u
        dt = time.perf_counter() - t0
        sample = "each step = one evaluation of the first %d patterns of the cfg2 workload (1/%d of it)" % (n_pat, spec["P"] // n_pat)
        used = cores
    else:
        n_loci = min(spec["M"], 128)
        trees, alns = make_locus_data(spec["M"], spec["n"], spec["L"], {"cfg3": 3, "cfg4": 4, "cfg5": 5}[wl], 0, n_loci)
        used = min(cores, 32)
        steps = min(steps, 10)
        for _ in range(warm):
            oracle_time_loci(trees, alns, used)
        dt, ups = 0.0, 0
        for _ in range(steps):
            d, u = oracle_time_loci(trees, alns, used)
            dt += d
            ups += u
        sample = "each step = the first %d of %d loci, split over %d host processes" % (n_loci, spec["M"], used)
    value = ups / dt
    print(json.dumps({
        "impl": "reference", "metric": "site-node partial updates/sec", "value": value, "unit": "updates/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "strong" if wl != "cfg2" else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s" % (wl, spec["desc"]), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "updates/s", "cores": used, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


if __name__ == "__main__":
    main()
