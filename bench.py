#!/usr/bin/env python
"""bench.py -- site-node partial updates/s of the sequence-likelihood hot path, at 1/2/4/8 B200.

    python bench.py --gpus 1 --steps K --warmup W            # cfg5 on one GPU (+ the cfg2 roofline run and the chain block)
    torchrun --nproc-per-node N bench.py --gpus N ...        # cfg5: 10,000 loci sharded over N GPUs
    python bench.py --impl reference ...                     # the reference's own CPU path on the host cores
    python bench.py --workload cfg2|cfg3|cfg4 ...            # one other BASELINE configuration as the headline

The headline workload is the SAME configuration at every N (BASELINE.json configs[4], "10,000 loci x 32 taxa x
2,000 sites sharded over 1/2/4/8 B200"), so the driver's 1 -> N efficiency is one curve on one workload.  One
"step" = one full log-likelihood evaluation of the workload with NEW branch lengths: P(t) of every branch,
every internal node recomputed, root reduction, and for N > 1 the collective that gives every rank all per-locus
values.  Prints ONE JSON line; DESIGN.md section 5 explains every field.
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
SEEDS = {"cfg3": 3, "cfg4": 4, "cfg5": 5}

WORKLOADS = {
    "cfg2": dict(M=1, n=50, P=1_000_000, desc="single 50-taxon gene tree, GTR, 1e6 unique site patterns"),
    "cfg3": dict(M=200, n=24, L=1000, desc="200 loci x 24 taxa x 1,000 sites, full logL eval"),
    "cfg4": dict(M=1000, n=16, L=500, desc="1,000 loci x 16 taxa x 500 sites"),
    "cfg5": dict(M=10_000, n=32, L=2000, desc="10,000 loci x 32 taxa x 2,000 sites sharded over N GPUs, one collective of the per-locus logL per evaluation"),
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


def make_locus_trees(M, n, seed, lo=0, hi=None, chunk=250):
    """Seeded random gene trees of loci [lo, hi) (per-chunk seeds: a shard sees the trees of the whole workload)."""
    from phynetpy_b200 import simulate as S
    hi = M if hi is None else hi
    trees = []
    for c0 in range((lo // chunk) * chunk, hi, chunk):
        c1 = min(M, c0 + chunk)
        par, bl = S.random_trees(c1 - c0, n, np.random.default_rng([seed, c0]))
        ft = S.flat_trees(par, bl)
        trees.extend(ft[i] for i in range(c1 - c0) if lo <= c0 + i < hi)
    return trees


def make_locus_data(M, n, L, seed, lo=0, hi=None, device=True):
    """Seeded synthetic loci [lo, hi): trees + (n, L) ASCII alignments evolved down them under the GTR model.  The
    sequences come from the counter-based simulator keyed by (seed, global locus index, site, node): on the device
    (K-s, pnb_simulate_alignments) for the GPU arm, from its NumPy twin (identical bytes, tests/test_gpu_chain.py) for
    the CPU arm -- any shard, on any number of ranks, on either side, sees the same data."""
    from phynetpy_b200 import simulate as S
    hi = M if hi is None else hi
    trees = make_locus_trees(M, n, seed, lo, hi)
    sim = S.simulate_alignments_device if device else S.simulate_alignments_counter
    alns = []
    for c0 in range(0, len(trees), 2500):      # bounded host / device memory per call
        alns.extend(sim(trees[c0:c0 + 2500], L, PI, RATES, seed, list(range(lo + c0, lo + min(c0 + 2500, len(trees))))))
    return trees, alns


def build_cfg2(P_target, n, seed, P):
    """50 taxa, exactly P_target unique patterns: simulate, compress on the GPU, keep the first P_target
    patterns with their weights (cfg 2)."""
    from phynetpy_b200 import simulate as S
    from phynetpy_b200._seq_likelihood import compress_alignment, _CODE_OF_BYTE
    rng = np.random.default_rng(seed)
    par, bl = S.random_trees(1, n, rng)
    ft = S.flat_trees(par, bl)[0]
    L = int(P_target * 1.06) + 64
    mat = S.simulate_alignments_device([ft], L, PI, RATES, seed, [0])[0]     # K-s: one thread per site
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
# The reference's own implementation of the path (baseline/_ref: the UNMODIFIED package, installed by
# baseline/install_reference.sh) when it is present, else the oracle port -- always on a bounded sample, with the
# per-locus objects built ONCE outside the clock, loci split over worker processes that live across steps.
def reference_available() -> bool:
    from baseline import ref_loader
    return ref_loader.available()


def _cpu_worker_main(conn, kind, payload):
    """One worker: build calculators + trees for its share once, then answer evaluation requests."""
    try:
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        from phynetpy_b200.tree import FlatTree, GeneTree
        items = []  # (evaluate(tree_index) -> float)
        if kind == "reference":
            from baseline import ref_loader
            ref_loader.load_reference()
            import phynetpy._seq_likelihood as RSL
            model = RSL.GTR(list(PI), list(RATES))
            for entry in payload:
                if entry[0] == "aln":      # (tag, taxa, matrix, [flat trees as (parent, blen, label)])
                    _, taxa, mat, trees = entry
                    calc = RSL.FelsensteinCalculator({t: mat[i].tobytes().decode() for i, t in enumerate(taxa)})
                    items.append((calc, [GeneTree(p.tolist(), [None if b != b else float(b) for b in bl.tolist()], lab)
                                         for p, bl, lab in trees], model, calc.n_patterns))
                else:                      # ("codes", codes, weights, tip_row, tree): compressed patterns (cfg2) -> the alignment
                    _, codes, weights, tip_row, (p, bl, lab) = entry      # they came from, column j repeated weights[j] times
                    cols = np.repeat(np.arange(codes.shape[1]), np.asarray(weights, dtype=np.int64))
                    # a code is the 4-bit set of allowed states (bit j = "ACGT"[j]); its IUPAC letter
                    assert int(codes.min()) >= 1 and int(codes.max()) <= 15
                    lut = np.frombuffer(b"-ACMGRSVTWYHKDBN", dtype=np.uint8)
                    aln = {lab[i]: lut[codes[int(r)][cols]].tobytes().decode() for i, r in enumerate(np.asarray(tip_row).tolist())
                           if r >= 0 and lab[i] is not None}
                    calc = RSL.FelsensteinCalculator(aln)
                    assert calc.n_patterns == codes.shape[1], (calc.n_patterns, codes.shape)
                    items.append((calc, [GeneTree(p.tolist(), [None if b != b else float(b) for b in bl.tolist()], lab)], model,
                                  calc.n_patterns))
        else:
            from oracle import felsenstein_oracle as O
            model = O.gtr(PI, RATES)
            for entry in payload:
                if entry[0] == "aln":
                    _, taxa, mat, trees = entry
                    locus = O.OracleLocus({t: mat[i].tobytes().decode() for i, t in enumerate(taxa)})
                    items.append((locus, [O.OracleTree(p.tolist(), [None if b != b else float(b) for b in bl.tolist()], lab)
                                          for p, bl, lab in trees], model, locus.n_patterns))
                else:                      # ("codes", codes, weights, tip_row, tree): already compressed (cfg2)
                    _, codes, weights, tip_row, (p, bl, lab) = entry
                    locus = O.OracleCodesLocus(codes, weights)
                    tree = O.OracleTree(p.tolist(), [None if b != b else float(b) for b in bl.tolist()], lab)
                    items.append((("codes", locus, tip_row), [tree], model, codes.shape[1]))

        def evaluate(obj, tree, model):
            if kind == "reference":
                return obj.log_likelihood(tree, model)
            from oracle import felsenstein_oracle as O
            if isinstance(obj, tuple):
                return obj[1].log_likelihood(obj[2], tree, model)
            return O.log_likelihood(obj, tree, model)
        conn.send(("ready", len(items)))
        while True:
            msg = conn.recv()
            if msg is None:
                break
            reps, which = msg           # which tree of every item (the trees differ by branch lengths)
            t0 = time.perf_counter()
            vals = None
            for _ in range(reps):
                vals = [evaluate(obj, trees[which % len(trees)], model) for obj, trees, model, _ in items]
            dt = (time.perf_counter() - t0) / reps
            ups = sum(npat * sum(1 for c in _n_children(trees[0]) if c > 0) for _, trees, _, npat in items)
            conn.send((dt, ups, vals))
    except Exception:
        import traceback
        conn.send(("error", traceback.format_exc()))


def _n_children(tree):
    parent = getattr(tree, "parent", None)
    if parent is None:
        parent = tree._parent
    n = len(parent)
    cnt = [0] * n
    for p in parent:
        if p >= 0:
            cnt[p] += 1
    return cnt


class CpuPath:
    """Worker processes holding the CPU implementation of the path for a bounded sample of a workload."""

    def __init__(self, entries, procs):
        import multiprocessing as mp
        self.kind = "reference" if reference_available() else "port"
        self.procs = max(1, min(procs, len(entries)))
        ctx = mp.get_context("spawn")
        self.workers = []
        for r in range(self.procs):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_cpu_worker_main, args=(b, self.kind, entries[r::self.procs]), daemon=True)
            p.start()
            self.workers.append((p, a))
        for _, a in self.workers:
            msg = a.recv()
            if msg[0] == "error":
                raise RuntimeError("CPU worker failed:\n" + msg[1])

    def step(self, reps=1, which=0):
        """One timed pass over the sample: max over workers of the mean evaluation time."""
        for _, a in self.workers:
            a.send((reps, which))
        res = [a.recv() for _, a in self.workers]
        for r in res:
            if r[0] == "error":
                raise RuntimeError("CPU worker failed:\n" + r[1])
        vals = [None] * sum(len(r[2]) for r in res)
        for k, r in enumerate(res):
            vals[k::self.procs] = r[2]
        return max(r[0] for r in res), sum(r[1] for r in res), vals

    def close(self):
        for p, a in self.workers:
            try:
                a.send(None)
            except Exception:
                pass
        for p, _ in self.workers:
            p.join(timeout=5)


def _tree_tuple(ft, scale=1.0):
    return (ft.parent, ft.blen * scale, ft.label)


def cpu_entries_loci(trees, alns, taxa, scales=(1.0,)):
    return [("aln", taxa, a, [_tree_tuple(t, s) for s in scales]) for t, a in zip(trees, alns)]


def describe_cpu(kind, procs, what):
    impl = ("the UNMODIFIED reference (baseline/_ref: phynetpy._seq_likelihood.FelsensteinCalculator.log_likelihood, "
            "NumPy/OpenBLAS)" if kind == "reference" else
            "oracle port (NumPy restatement of the reference's _postorder_partials; baseline/_ref is absent)")
    return "%s on %s, calculators built once outside the clock, loci split over %d worker process(es)" % (impl, what, procs)


# --------------------------------------------------------------------------- device workload
class Workload:
    """A resident workload on this rank: loci in HBM, a resident batch, branch-length sets on host and device."""

    def __init__(self, P, ctx, model, calcs, trees, n_sets):
        import torch
        from phynetpy_b200._seq_likelihood import _model_handle
        from phynetpy_b200.device import ResidentBatch, pack_trees
        self.P, self.ctx, self.model, self.calcs, self.trees = P, ctx, model, calcs, trees
        for c in calcs:
            c._locus_handle()
        self.handles = np.array([c._locus.value for c in calcs], dtype=np.uint64)
        self.node_off, self.parent, self.blen, self.tip_row = pack_trees(trees, [c._row_of for c in calcs])
        self.mh = _model_handle(model, ctx)
        self.batch = ResidentBatch(ctx, self.mh, self.handles, self.node_off, self.parent, self.tip_row)
        self.score_batch = None
        # every step scores the same topologies with NEW branch lengths (a height-rescaling proposal)
        self.h_blens = [torch.from_numpy(np.ascontiguousarray(self.blen * (1.0 + 0.002 * k))).pin_memory() for k in range(n_sets)]
        self.d_blens = [t.cuda() for t in self.h_blens]
        self.n_local = len(calcs)
        self.patterns_local = int(sum(c.n_patterns for c in calcs))
        self.updates_local = self.batch.info["updates"]

    def make_score_batch(self):
        from phynetpy_b200 import _lib
        from phynetpy_b200.device import ResidentBatch
        if self.score_batch is None:
            self.score_batch = ResidentBatch(self.ctx, self.mh, self.handles, self.node_off, self.parent, self.tip_row, _lib.EVAL_NOSTORE)
        return self.score_batch

    def close(self):
        for b in (self.batch, self.score_batch):
            if b is not None:
                b.close()
        for c in self.calcs:
            c.close()


def measure(args, wl_name, W, plan, stream, rank, world, clock_sampler=None):
    """value / e2e / per-kernel numbers of one resident workload (this rank's shard), timed as the contract says."""
    import torch
    import torch.distributed as dist
    ctx = W.ctx
    M = plan["M"]
    lo, per = plan["lo"], plan["per"]
    vec_len = per * world
    # Every rank ends with all per-locus values.  Preferred: the vector lives in symmetric memory (every peer's copy is
    # mapped here) and the pruning kernel itself stores each finished locus into every peer's vector -- the collective is
    # fused into the kernel's epilogue, followed by one device-side barrier.  Otherwise: in-place ncclAllGather.
    sym = None
    if world > 1 and not args.skip_collective and args.collective == "p2p":
        from phynetpy_b200.distributed import ShardPlan
        sp = ShardPlan(M, rank, world, [(r * per, min(M, (r + 1) * per)) for r in range(world)])
        sym = sp.symmetric_vector()
    if sym is not None:
        # two copies used alternately by consecutive steps: a fast rank's next kernel never stores into a vector that a slower
        # rank is still copying to its host (it cannot get two steps ahead: it waits in the barrier in between)
        sym_t, sym_hdl = sym
        vecs = [sym_t[:vec_len], sym_t[vec_len:2 * vec_len]]
        peer_sets = [[sym_hdl.buffer_ptrs[r] + 8 * (c * vec_len + lo) for r in range(world) if r != rank] for c in range(2)]
        bar_args = sp.peer_barrier_args(sym_hdl)     # the barrier is issued by the kernel's last block (no launch after it)
    else:
        vecs, sym_hdl, peer_sets = [torch.zeros(vec_len, dtype=torch.float64, device="cuda")] * 2, None, [[], []]
    out_host = torch.zeros(vec_len, dtype=torch.float64).pin_memory()
    cur = {"vec": vecs[0]}
    steps = args.steps
    warm = max(args.warmup, 3)
    step_no = [0]
    n_sets = len(W.d_blens)
    mode = ["p2p" if sym is not None else "nccl"]
    inplace_ok = [True]
    epoch = [0]

    def set_mode(m):
        mode[0] = m
        if m != "p2p":
            W.batch.set_peer_outputs([])

    def begin_step():
        """-> (index of the branch-length set, this rank's slice of the output vector of this step)"""
        k = step_no[0]
        step_no[0] += 1
        c = k % 2 if mode[0] == "p2p" else 0
        cur["vec"] = vecs[c]
        if mode[0] == "p2p":
            W.batch.set_peer_outputs(peer_sets[c])
            if not args.torch_barrier:
                epoch[0] += 1
                W.batch.set_peer_barrier(*bar_args, epoch[0])
        return k % n_sets, vecs[c].data_ptr() + 8 * lo

    def collective():
        out_vec = cur["vec"]
        if world > 1 and not args.skip_collective:
            if mode[0] == "p2p":
                # every peer's stores must have landed before anyone reads its vector: the kernel's last block has already
                # exchanged epochs with the peers (pnb_batch_set_peer_barrier); --torch-barrier: a barrier launch instead
                if args.torch_barrier:
                    sym_hdl.barrier()
            elif inplace_ok[0]:
                # in place: this rank's slice of the M-vector IS the send buffer (equal slices by construction)
                dist.all_gather_into_tensor(out_vec, out_vec[lo:lo + per])
            else:
                dist.all_gather_into_tensor(out_vec, out_vec[lo:lo + per].clone())

    def step_device():
        """inputs resident in HBM (lengths included); logL stays on the device; collective when sharded"""
        k, out_local_ptr = begin_step()
        with torch.cuda.stream(stream):
            W.batch.eval(W.mh, W.d_blens[k].data_ptr(), out_device_ptr=out_local_ptr)
            collective()

    def step_e2e():
        """the call a user makes: host branch lengths (pinned) in, host log-likelihoods of ALL loci out"""
        k, out_local_ptr = begin_step()
        with torch.cuda.stream(stream):
            W.batch.eval(W.mh, W.h_blens[k].numpy(), out_device_ptr=out_local_ptr)
            collective()
            out_host.copy_(cur["vec"], non_blocking=True)
        stream.synchronize()
        return out_host.numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_device_loop(n=None):
        n = steps if n is None else n
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        barrier()
        step_no[0] = 0
        ev0.record(stream)
        for _ in range(n):
            step_device()
        ev1.record(stream)
        barrier()
        return ev0.elapsed_time(ev1)

    # ---- warm-up; the three ways of combining the ranks' values must agree bit for bit ----------------------
    set_mode("nccl")
    inplace_ok[0] = False
    for _ in range(warm):
        step_device()
    barrier()
    nccl_ms = p2p_ms = None
    if world > 1 and not args.skip_collective:
        ref_vec = cur["vec"].clone()
        inplace_ok[0] = True
        step_no[0] -= 1
        step_device()
        barrier()
        flags_ok = [bool(torch.equal(ref_vec, cur["vec"])), True]
        if sym is not None:
            set_mode("p2p")
            vecs[0].zero_(); vecs[1].zero_()
            barrier()
            step_no[0] -= 1
            step_device()
            barrier()
            flags_ok[1] = bool(torch.equal(ref_vec, cur["vec"]))
        t = torch.tensor([1.0 if f else 0.0 for f in flags_ok], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        inplace_ok[0] = bool(t[0].item() > 0.5)
        if sym is not None and t[1].item() < 0.5:
            raise RuntimeError("fused peer stores disagree with the NCCL all-gather")
        if sym is not None:
            # both variants of the same step timed the same way; the faster one (agreed over ranks) runs the timed region
            # (two rounds each, alternating, best of the two: long kernels run under the power cap, and whichever variant is
            # timed later in a run would otherwise be charged for the warmer GPU)
            # ... and short -- 2 x 6 steps per variant: the trials only choose the variant, the timed region below is the
            # measurement, and every step run before it costs the timed region clock under the power cap)
            nccl_ms = p2p_ms = float("inf")
            n_trial = max(3, min(steps, 6))
            for _round in range(2):
                set_mode("nccl")
                for _ in range(2):
                    step_device()
                nccl_ms = min(nccl_ms, timed_device_loop(n_trial) / n_trial)
                set_mode("p2p")
                for _ in range(2):
                    step_device()
                p2p_ms = min(p2p_ms, timed_device_loop(n_trial) / n_trial)
            tt = torch.tensor([nccl_ms, p2p_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            nccl_ms, p2p_ms = float(tt[0]), float(tt[1])
            if args.collective_auto and nccl_ms < p2p_ms:
                set_mode("nccl")
            for _ in range(warm):
                step_device()
            barrier()
    st = ctx.last_stats()
    launches_per_step = st["launches"]

    def timed_regions():
        # ---- timed: device-resident -----------------------------------------------------------------------
        if clock_sampler is not None:
            clock_sampler.start()
        ctx.set_profiling(False)
        elapsed_ms = timed_device_loop()
        logL_device = cur["vec"].cpu().numpy().copy()

        # ---- per-kernel durations: the same steps with CUDA events around each launch (library-side, same stream) ----
        ctx.set_profiling(True)
        prune_ms, pm_ms, call_ms = [], [], []
        step_no[0] = 0
        for _ in range(steps):
            step_device()
            tm = ctx.last_timings()
            prune_ms.append(tm["prune_ms"]); pm_ms.append(tm["pmatrix_ms"]); call_ms.append(tm["call_ms"])
        barrier()
        ctx.set_profiling(False)

        # ---- timed: end to end with host buffers ----------------------------------------------------------------
        for _ in range(2):
            step_e2e()
        barrier()
        step_no[0] = 0
        t0 = time.perf_counter()
        for _ in range(steps):
            vals_e2e = step_e2e()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        st_e2e = ctx.last_stats()
        logL_e2e = vals_e2e.copy()
        barrier()
        clocks = clock_sampler.stop() if clock_sampler is not None else None

        return elapsed_ms, logL_device, prune_ms, pm_ms, call_ms, e2e_ms, st_e2e, logL_e2e, clocks

    elapsed_ms, logL_device, prune_ms, pm_ms, call_ms, e2e_ms, st_e2e, logL_e2e, clocks = timed_regions()
    fell_back = False
    if world > 1 and mode[0] == "p2p":
        # the kernel-issued barrier gives up on a peer after ~4 s instead of hanging (and says so): a run in which that
        # happened anywhere is measured again with the NCCL all-gather
        tflag = torch.tensor([1.0 if W.batch.barrier_timed_out() else 0.0], device="cuda")
        dist.all_reduce(tflag, op=dist.ReduceOp.MAX)
        if tflag.item() > 0.5:
            fell_back = True
            set_mode("nccl")
            for _ in range(warm):
                step_device()
            barrier()
            elapsed_ms, logL_device, prune_ms, pm_ms, call_ms, e2e_ms, st_e2e, logL_e2e, clocks = timed_regions()

    # ---- reduce over ranks ------------------------------------------------------------------------------
    t = torch.tensor([elapsed_ms, e2e_ms], dtype=torch.float64, device="cuda")
    u = torch.tensor([float(W.updates_local), float(W.patterns_local)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return {
        "elapsed_ms": float(t[0]), "e2e_ms": float(t[1]), "updates_total": float(u[0]), "patterns_total": int(u[1]),
        "kernel_ms": float(np.mean(prune_ms)), "pmatrix_ms": float(np.mean(pm_ms)), "call_ms": float(np.mean(call_ms)),
        "launches_per_step": int(launches_per_step), "h2d": int(st_e2e["h2d_bytes"]),
        "d2h": int(vec_len * 8), "logL_device": logL_device[:M], "logL_e2e": logL_e2e[:M], "clocks": clocks,
        "inplace_allgather": bool(inplace_ok[0]) if world > 1 else None,
        "collective_mode": None if world == 1 else mode[0], "nccl_ms_per_step": nccl_ms, "p2p_ms_per_step": p2p_ms,
        "barrier_timed_out": bool(fell_back),
    }


def score_only_leg(args, W, stream):
    """PNB_EVAL_NOSTORE: the same evaluation with no partial written (candidate scoring)."""
    import torch
    sb = W.make_score_batch()
    out = torch.zeros(W.n_local, dtype=torch.float64, device="cuda")
    W.ctx.set_profiling(True)
    ks = []
    for k in range(3 + args.steps):
        with torch.cuda.stream(stream):
            sb.eval(W.mh, W.d_blens[k % len(W.d_blens)].data_ptr(), out_device_ptr=out.data_ptr())
        tm = W.ctx.last_timings()
        if k >= 3:
            ks.append(tm["prune_ms"])
    # same lengths as the last timed store-mode step for the equality check
    with torch.cuda.stream(stream):
        sb.eval(W.mh, W.d_blens[(args.steps - 1) % len(W.d_blens)].data_ptr(), out_device_ptr=out.data_ptr())
    stream.synchronize()
    W.ctx.set_profiling(False)
    return float(np.mean(ks)), out.cpu().numpy()


def roofline_block(W, kernel_ms, pm_ms, call_ms, traffic_ncu_file=None):
    peak, peak_src = peak_hbm()
    tr = W.batch.traffic
    compulsory = tr["partials_written"] + tr["tip_codes"] + 8 * sum(c.n_patterns for c in W.calcs)
    achieved_real = compulsory / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else None
    achieved_model = W.updates_local * ALGO_BYTES_PER_UPDATE / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else None
    out = {
        "bound": "hbm", "kernel": "pnb_prune_kernel",
        "achieved": achieved_real, "peak": peak, "unit": "GB/s",
        "frac": (achieved_real / peak) if achieved_real else None,
        "traffic": compulsory,
        "traffic_source": "counted live from the compiled op programs of this run (pnb_batch_traffic): node partials the kernel "
                          "writes (stored nodes x padded patterns x 32 B) + tip-code rows + weights; re-reads of siblings "
                          "written in the same launch (%d B) and per-tile program / P staging (%d B) are L2 traffic and not counted"
                          % (tr["partials_reread"], tr["weights_programs_pmatrices"]),
        "bytes_definition": "achieved / frac use the bytes THIS kernel has to move (a thread carries its patterns through the whole "
                            "tree: tips are 1-byte codes, a just-computed child stays in registers, tip-only subtrees are never "
                            "stored); model_96B is SURVEY section 8d's streaming model (2 x 32 B read + 32 B written per update)",
        "model_96B": {"achieved": achieved_model, "frac": (achieved_model / peak) if achieved_model else None,
                      "algorithmic_bytes_per_update": ALGO_BYTES_PER_UPDATE},
        "peak_source": peak_src,
        "kernel_ms": kernel_ms, "pmatrix_kernel_ms": pm_ms, "host_call_ms": call_ms,
    }
    if traffic_ncu_file and os.path.exists(os.path.join(ROOT, traffic_ncu_file)):
        try:
            d = json.load(open(os.path.join(ROOT, traffic_ncu_file)))
            out["traffic_ncu"] = {"dram_bytes_per_launch": d.get("dram_bytes_per_launch"), "file": traffic_ncu_file,
                                  "note": "ncu --set full capture of this kernel at the recorded commit (a committed file, not this run)"}
        except Exception:
            pass
    return out


# --------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (smoke runs only; reported)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="N=1: skip the cfg2 roofline block and the chain block")
    ap.add_argument("--skip-collective", action="store_true", help="diagnostic only: time the sharded step without the collective")
    ap.add_argument("--no-collective-auto", dest="collective_auto", action="store_false",
                    help="keep the fused peer stores even when the NCCL all-gather timed faster in the trial")
    ap.add_argument("--torch-barrier", action="store_true",
                    help="fused peer stores: order them with a separate symmetric-memory barrier launch instead of the kernel's own barrier")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: fused peer stores from the kernel (falls back to NCCL when symmetric memory is unavailable) or ncclAllGather")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = args.workload
    extras = False
    if wl == "auto":
        wl = "cfg5"
        extras = world == 1 and not args.no_extras and args.scale == 1.0
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

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    os.environ["PHYNET_B200_DEVICE"] = str(local_rank)
    ctx = P.default_context(local_rank)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    model = P.GTR(PI, RATES)
    n_sets = max(min(args.steps, 8), 1)

    cfg2_block = chain = None
    if extras:
        cfg2_block = run_cfg2_block(args, P, ctx, model, stream, n_sets)
        chain = run_chain_block(args, P, ctx, model)

    # ---- build the resident headline workload ---------------------------------------------------------
    t_setup0 = time.perf_counter()
    setup = {}
    cpu_entries = None
    if wl == "cfg2":
        if world > 1:
            raise SystemExit("cfg2 is a single-locus, single-GPU workload (replicas only); use cfg5 for N>1")
        fc, ft, setup = build_cfg2(spec["P"], spec["n"], 2, P)
        calcs, trees = [fc], [ft]
        M, lo, per = 1, 0, 1
        n_pat = min(200_000, fc.n_patterns)
        cpu_entries = [("codes", np.ascontiguousarray(fc._codes[:, :n_pat]), fc._int_weights[:n_pat].copy(),
                        ft.tip_rows(fc._row_of).tolist(), _tree_tuple(ft))]
        cpu_what = "the first %d of the %d patterns" % (n_pat, fc.n_patterns)
    else:
        M = spec["M"]
        per = (M + world - 1) // world              # equal contiguous slices (the loci of one workload do equal work)
        lo, hi = rank * per, min(M, (rank + 1) * per)
        trees, alns = make_locus_data(M, spec["n"], spec["L"], SEEDS[wl], lo, hi)
        taxa = ["t%d" % i for i in range(spec["n"])]
        calcs = [P.FelsensteinCalculator.from_matrix(taxa, a, device_compress=False) for a in alns]
        n_cpu = min(len(trees), 512)
        cpu_entries = cpu_entries_loci(trees[:n_cpu], alns[:n_cpu], taxa)
        cpu_what = "the first %d of the %d loci" % (n_cpu, M)
        del alns
    W = Workload(P, ctx, model, calcs, trees, n_sets)
    setup["setup_s"] = round(time.perf_counter() - t_setup0, 2)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    m = measure(args, wl, W, {"M": M, "lo": lo, "per": per}, stream, rank, world, sampler)
    score_ms = score_vals = None
    if world == 1:
        score_ms, score_vals = score_only_leg(args, W, stream)

    if rank == 0:
        steps = args.steps
        ms_per_step = m["elapsed_ms"] / steps
        value = m["updates_total"] * steps / (m["elapsed_ms"] * 1e-3)
        e2e_value = m["updates_total"] * steps / (m["e2e_ms"] * 1e-3)
        total_logL = float(np.sum(m["logL_device"]))
        result = {
            "metric": "site-node partial updates/sec",
            "value": value,
            "unit": "updates/s",
            "n_gpus": world,
            "steps": steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "strong" if wl != "cfg2" else "weak",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": "%s: %s" % (wl, spec["desc"]),
                "loci": M, "taxa": spec["n"], "site_patterns_total": m["patterns_total"],
                "updates_per_step": int(m["updates_total"]),
                "model": "GTR pi=(.1,.2,.3,.4) rates=(1,2.5,.7,1.3,3.1,.9)",
                "step": "full logL evaluation of a fresh proposal (same topologies, new branch lengths every step): P(t) of every "
                        "branch + every internal node recomputed + root reduction"
                        + (" + every rank ends with all %d per-locus values (see collective)" % M if world > 1 else "")
                        + "; op programs resident in HBM (pnb_batch_*), compiled once per topology",
                "parallelism": "loci sharded over %d GPU(s) in equal contiguous slices, one process per GPU" % world,
                "collective": None if world == 1 else (
                    "fused into the pruning kernel: the block that completes a locus stores its logL into every peer's vector "
                    "(symmetric memory, NVLink peer stores) and the block that completes the LAST locus exchanges an epoch flag with "
                    "every peer before the kernel ends (%s); checked bit-equal to the NCCL all-gather in this run"
                    % ("--torch-barrier: a separate barrier launch instead" if args.torch_barrier else "no launch follows the kernel")
                    if m["collective_mode"] == "p2p" else
                    ("ncclAllGather in place on the kernel's stream" if m["inplace_allgather"] else "ncclAllGather (out of place)")),
                "barrier_timed_out": m["barrier_timed_out"],
                "collective_trials_ms_per_step": {"nccl_all_gather_in_place": m["nccl_ms_per_step"],
                                                  "fused_peer_stores": m["p2p_ms_per_step"]},
                "l2": "working set per step (%.2f GB of partials written per GPU) exceeds the 126 MB L2; no flush needed"
                      % (W.batch.traffic["partials_written"] / 1e9),
                "scale": args.scale, "diagnostic_skip_collective": bool(args.skip_collective),
                **setup,
            },
            "e2e": {
                "value": e2e_value, "unit": "updates/s", "ms_per_step": m["e2e_ms"] / steps,
                "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                "what": "pnb_batch_eval through the C ABI with HOST buffers: branch lengths of every tree in pinned host memory -> "
                        "chunked H2D overlapped with pruning -> P(t) + pruning kernels -> "
                        + ("peer stores / all-gather -> " if world > 1 else "") + "per-locus logL of ALL loci copied D2H into pinned host memory "
                        "(per-rank bytes); alignments resident since construction",
            },
            "gpu_launches": int(m["launches_per_step"] * steps),
            "kernels_per_step": {"pnb_pmatrix_gather_kernel": 1, "pnb_prune_kernel": 1},
            "roofline": roofline_block(W, m["kernel_ms"], m["pmatrix_ms"], m["call_ms"],
                                       "profiles/traffic_%s_r2.json" % wl),
            "score_only": None if score_ms is None else {
                "what": "PNB_EVAL_NOSTORE: every node recomputed, no partial written to HBM (candidate scoring)",
                "kernel_ms": score_ms, "updates_per_s_kernel": W.updates_local / (score_ms * 1e-3),
                "logL_equal_to_store_mode": bool(np.array_equal(score_vals, m["logL_device"][:len(score_vals)])),
            },
            "cfg2": cfg2_block,
            "chain": chain,
            "clocks": m["clocks"],
            "logL": total_logL,
            "logL_e2e": float(np.sum(m["logL_e2e"])),
            "logL_e2e_equals_device_bitwise": bool(np.array_equal(m["logL_e2e"], m["logL_device"])),
        }
        if not args.no_cpu_baseline and world == 1:
            result["cpu_baseline"], result["parity"] = cpu_baseline(cpu_entries, cpu_what, W, m)
        print(json.dumps(result))
    W.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(entries, what, W, m):
    """The reference (or its port) on a bounded sample of the SAME inputs on this box's host cores, and the
    in-run parity check of the GPU values against it."""
    cores = os.cpu_count() or 1
    procs = min(cores, 32)
    cp = CpuPath(entries, procs)
    try:
        cp.step(1)                                   # warm-up
        reps = 3 if entries[0][0] == "codes" else 5
        best, ups, vals = None, 0, None
        for _ in range(3):
            dt, ups, vals = cp.step(reps)
            best = dt if best is None else min(best, dt)
    finally:
        cp.close()
    base = {"value": ups / best, "unit": "updates/s", "cores": cp.procs, "kind": cp.kind,
            "sample": describe_cpu(cp.kind, cp.procs, what) + "; best of 3 timings of %d evaluations each" % reps,
            "seconds_per_eval_sample": best}
    # parity: the GPU's values for the same loci at the same (unscaled) branch lengths
    if entries[0][0] == "codes":
        n_pat = entries[0][1].shape[1]
        sub = W.P.FelsensteinCalculator.from_codes(W.calcs[0].taxa, entries[0][1], entries[0][2])
        got = [sub.log_likelihood(W.trees[0], W.model)]
        sub.close()
        label = "first %d patterns" % n_pat
    else:
        n = len(entries)
        got = W.batch.eval(W.mh, np.ascontiguousarray(W.blen))[:n].tolist()
        label = "first %d loci" % n
    err = max(abs(g - v) / max(1.0, abs(v)) for g, v in zip(got, vals))
    parity = {"max_rel_err_vs_cpu": err, "cpu_kind": cp.kind, "sample": label,
              "cpu_logL_sample_sum": float(sum(vals)), "gpu_logL_sample_sum": float(sum(got))}
    return base, parity


# --------------------------------------------------------------------------- N=1 extras
def run_cfg2_block(args, P, ctx, model, stream, n_sets):
    """BASELINE configs[1], the single-GPU roofline run: one 50-taxon gene tree, 1e6 unique site patterns."""
    spec = WORKLOADS["cfg2"]
    fc, ft, setup = build_cfg2(spec["P"], spec["n"], 2, P)
    W = Workload(P, ctx, model, [fc], [ft], n_sets)
    m = measure(args, "cfg2", W, {"M": 1, "lo": 0, "per": 1}, stream, 0, 1, None)
    score_ms, score_vals = score_only_leg(args, W, stream)
    steps = args.steps
    # the same call one level up: the reference-shaped Python facade with a Network-like tree OBJECT
    gt = P.GeneTree.from_flat(ft)
    fc.log_likelihood(gt, model, full=True)
    t0 = time.perf_counter()
    nrep = min(steps, 10)
    for _ in range(nrep):
        fc.log_likelihood(gt, model, full=True)
    facade_ms = (time.perf_counter() - t0) * 1e3 / nrep
    out = {
        "workload": "cfg2: " + spec["desc"],
        "value": m["updates_total"] * steps / (m["elapsed_ms"] * 1e-3), "unit": "updates/s",
        "ms_per_step": m["elapsed_ms"] / steps,
        "e2e": {"value": m["updates_total"] * steps / (m["e2e_ms"] * 1e-3), "ms_per_step": m["e2e_ms"] / steps,
                "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                "python_facade_ms_per_step": facade_ms,
                "python_facade_what": "FelsensteinCalculator.log_likelihood(tree object, model, full=True): per-call marshalling "
                                      "of a Network-like object and tree compilation included"},
        "roofline": roofline_block(W, m["kernel_ms"], m["pmatrix_ms"], m["call_ms"], "profiles/traffic_cfg2_r2.json"),
        "score_only": {"kernel_ms": score_ms, "updates_per_s_kernel": W.updates_local / (score_ms * 1e-3),
                       "logL_equal_to_store_mode": bool(np.array_equal(score_vals, m["logL_device"][:1]))},
        "logL": float(m["logL_device"][0]),
        **setup,
    }
    if not args.no_cpu_baseline:
        n_pat = min(200_000, fc.n_patterns)
        entries = [("codes", np.ascontiguousarray(fc._codes[:, :n_pat]), fc._int_weights[:n_pat].copy(),
                    ft.tip_rows(fc._row_of).tolist(), _tree_tuple(ft))]
        out["cpu_baseline"], out["parity"] = cpu_baseline(entries, "the first %d of the %d patterns" % (n_pat, fc.n_patterns), W, m)
    W.close()
    return out


def run_chain_block(args, P, ctx, model):
    """The second half of the metric -- MCMC_SEQ logL evaluations / s -- on BASELINE configs[2] (cfg3: full logL of
    every locus per proposal) and configs[3] (cfg4: chain steps with dirty-path recompute), each with the CPU path
    timed on the same inputs and a bitwise cached == recomputed flag.  Bounded to a few seconds each."""
    out = {}
    out["cfg3_full_eval"] = chain_cfg3(args, P, ctx, model)
    out["cfg4_dirty_chain"], out["cfg4_joint_chain"] = chain_cfg4(args, P, ctx, model)
    if not args.no_cpu_baseline:
        try:
            out["reference_mcmc_seq"] = chain_dropin(args)
        except Exception:      # the reference's own sampler is the one leg here that is not our code: say what happened, keep the line
            import traceback
            from phynetpy_b200 import dropin
            dropin.uninstall()
            out["reference_mcmc_seq"] = {"error": traceback.format_exc()[-1500:]}
    return out


def chain_cfg3(args, P, ctx, model):
    """cfg3: 200 loci x 24 taxa x 1,000 sites; every proposal re-scores EVERY locus (reference loop
    src/_mcmc_seq.py:746-750 with every _felsen entry None), through the engine seam and through the resident batch."""
    spec = WORKLOADS["cfg3"]
    trees, alns = make_locus_data(spec["M"], spec["n"], spec["L"], SEEDS["cfg3"])
    taxa = ["t%d" % i for i in range(spec["n"])]
    calcs = [P.FelsensteinCalculator.from_matrix(taxa, a, device_compress=False) for a in alns]
    eng = P.SeqLikelihoodEngine(calcs, None, model, context=ctx, resident_batches=True)
    scales = [1.0 + 0.002 * k for k in range(8)]
    props = [[t.with_lengths(t.blen * s) for t in trees] for s in scales]
    M = len(trees)

    def one(k):
        for i in range(M):
            eng.invalidate_locus(i)
        return eng.log_likelihood(props[k % 8], None, 0.0)
    for k in range(3):
        one(k)
    n = 100
    t0 = time.perf_counter()
    for k in range(n):
        total = one(k)
    seam_s = (time.perf_counter() - t0) / n
    eng._commit_pending()
    cached = list(eng._felsen)
    for i in range(M):
        eng.invalidate_locus(i)
    eng.invalidate_device_cache()
    fresh_total = eng.log_likelihood(props[(n - 1) % 8], None, 0.0)
    eng._commit_pending()
    # the same evaluation through the resident batch (lengths only)
    import torch
    W = Workload(P, ctx, model, calcs, trees, 8)
    for k in range(3):
        W.batch.eval(W.mh, W.h_blens[k].numpy())
    t0 = time.perf_counter()
    for k in range(n):
        vals = W.batch.eval(W.mh, W.h_blens[k % 8].numpy())
    batch_s = (time.perf_counter() - t0) / n
    res = {
        "what": "full logL of all 200 loci per proposal, host trees / lengths in -> host values out, Python loop included",
        "evals_per_s_engine_seam": 1.0 / seam_s, "evals_per_s_resident_batch": 1.0 / batch_s,
        "updates_per_eval": int(W.updates_local),
        "cached_equals_recomputed_bitwise": bool(cached == list(eng._felsen) and fresh_total == total),
        "batch_equals_engine_bitwise": bool(np.array_equal(vals, np.array(cached))),
    }
    if not args.no_cpu_baseline:
        cp = CpuPath(cpu_entries_loci(trees, alns, taxa, scales=(scales[(n - 1) % 8],)), min(os.cpu_count() or 1, 32))
        try:
            cp.step(1)
            dt, _, cvals = cp.step(3)
        finally:
            cp.close()
        res["cpu"] = {"evals_per_s": 1.0 / dt, "kind": cp.kind, "cores": cp.procs,
                      "sample": describe_cpu(cp.kind, cp.procs, "all 200 loci (the whole evaluation)"),
                      "max_rel_err_gpu_vs_cpu": max(abs(g - v) / abs(v) for g, v in zip(cached, cvals))}
    W.batch.close()
    for c in calcs:
        c.close()
    return res


def chain_dropin(args, n_loci=30, n_taxa=10, n_sites=400, n_iter=120):
    """The reference's OWN sampler -- ``MCMC_SEQ(...).search()``, unmodified Python from ``baseline/_ref`` -- timed twice on
    the same data and seed: as it is (every likelihood on the host), and with ``phynetpy_b200.dropin.install()`` (every
    Felsenstein pruning, every MSNC density and the placement scans of the informed network moves on the device).  What
    a user of the reference gets by switching; iterations / s include all of the reference's Python (operators on
    Network objects, priors, bookkeeping).  The chain is not reproducible from its seed (set iteration over objects
    hashed by address), so the two arms walk different trajectories of the same sampler."""
    if not reference_available():
        return {"unavailable": "baseline/_ref is not installed"}
    from baseline import ref_loader
    from phynetpy_b200 import simulate as S, dropin
    ref = ref_loader.load_reference()
    import phynetpy._mcmc_seq as mcmc
    rng = np.random.default_rng(5)
    net = S.random_species_network(n_taxa, 0, rng, height=0.005 * (n_taxa - 1))
    trees, _ = S.simulate_gene_trees_msnc(net, 1, 0.02, rng, n_loci)
    aln = S.simulate_alignments(np.stack([t.parent for t in trees]), np.stack([t.blen for t in trees]), n_sites,
                                S.DEFAULT_PI, S.DEFAULT_RATES, rng)
    loci = [S.to_alignment_dict(aln[i]) for i in range(n_loci)]
    mapping = {t: [t] for t in loci[0].keys()}

    def arm():
        inf = mcmc.MCMC_SEQ(loci, mapping, priors=mcmc.MCMCSeqPriors(max_reticulations=1))
        s0 = inf.score()
        t0 = time.perf_counter()
        res = inf.search(num_iter=n_iter, burn_in=10, sample_freq=10, seed=7)
        return s0, time.perf_counter() - t0, res

    dropin.install(ref)
    try:
        arm()                                  # warm-up: contexts, handles, first launches
        s_gpu, t_gpu, r_gpu = arm()
    finally:
        dropin.uninstall()
    s_cpu, t_cpu, r_cpu = arm()
    return {
        "what": "the reference's own MCMC_SEQ.search(num_iter=%d) on %d loci x %d taxa x %d sites, max 1 reticulation, "
                "default operator mixture (gene-tree moves, theta / height moves, informed add / delete of reticulations)"
                % (n_iter, n_loci, n_taxa, n_sites),
        "iterations_per_s_with_drop_in": n_iter / t_gpu, "iterations_per_s_reference_alone": n_iter / t_cpu,
        "seconds_with_drop_in": t_gpu, "seconds_reference_alone": t_cpu,
        "initial_score_rel_diff": abs(s_gpu - s_cpu) / max(1.0, abs(s_cpu)),
        "acceptance_rates": [r_gpu.acceptance_rate, r_cpu.acceptance_rate],
        "cores_reference": 1,
    }


def chain_cfg4(args, P, ctx, model):
    """cfg4: 1,000 loci x 16 taxa x 500 sites as MCMC chains.  (a) Felsenstein only: single-locus node-height
    proposals through the engine seam (dirty-path recompute, commit on accept / rollback on reject).  (b) Both
    factors of the MCMC_SEQ likelihood (reference src/_mcmc_seq.py:745-761): the reference's gene-tree operators on
    flat arrays + a theta move every 10th step, on a 1-reticulation network."""
    from phynetpy_b200 import simulate as S, moves as MV
    from phynetpy_b200.candidates import node_heights
    spec = WORKLOADS["cfg4"]
    n_loci, n_species, L = spec["M"], spec["n"], spec["L"]
    rng = np.random.default_rng(45)
    theta = 0.02
    net = S.random_species_network(n_species, 1, rng, height=0.005 * (n_species - 1))
    trees, species_of = S.simulate_gene_trees_msnc(net, 1, theta, rng, n_loci)
    aln = S.simulate_alignments(np.stack([t.parent for t in trees]), np.stack([t.blen for t in trees]), L,
                                S.DEFAULT_PI, S.DEFAULT_RATES, rng)
    loci = [S.to_alignment_dict(aln[i]) for i in range(n_loci)]
    taxa = list(loci[0].keys())

    # ---- (a) dirty-path chain, Felsenstein factor only -------------------------------------------------
    eng = P.SeqLikelihoodEngine(loci, species_of, model, context=ctx)
    cur = list(trees)
    heights = [node_heights(t) for t in cur]
    total = eng.log_likelihood(cur, None, 0.0)
    n_steps = 1500
    accepted = recomputed = 0
    lat = []
    touched = []
    t0 = time.perf_counter()
    for step in range(n_steps):
        i = int(rng.integers(0, n_loci))
        mv = MV.slide_height_move(cur[i], heights[i], rng)
        if mv is None:
            continue
        snap = eng.clone_caches()
        prop = list(cur)
        prop[i] = mv[0]
        eng.invalidate_locus(i)
        s0 = time.perf_counter()
        new_total = eng.log_likelihood(prop, None, 0.0)
        lat.append(time.perf_counter() - s0)
        recomputed += eng.last_stats["nodes_recomputed"]
        touched.append(i)
        if math.log(rng.random()) < (new_total - total) + mv[2]:
            cur, total = prop, new_total
            heights[i] = mv[1]
            accepted += 1
        else:
            eng.restore_caches(snap)
    wall = time.perf_counter() - t0
    eng._commit_pending()
    inc = list(eng._felsen)
    for i in range(n_loci):
        eng.invalidate_locus(i)
    eng.invalidate_device_cache()
    full = eng.log_likelihood(cur, None, 0.0)
    eng._commit_pending()
    dirty = {
        "what": "single-locus node-height proposals (the reference's op_gene_node_height on flat arrays) through "
                "SeqLikelihoodEngine: dirty-path recompute, commit on accept / rollback on reject, Python loop included",
        "steps": n_steps, "steps_per_s": len(lat) / wall, "accept_rate": accepted / max(len(lat), 1),
        "mean_nodes_recomputed": recomputed / max(len(lat), 1),
        "median_eval_latency_us": float(np.median(lat) * 1e6),
        "cached_equals_recomputed_bitwise": bool(inc == list(eng._felsen) and full == total),
    }
    if not args.no_cpu_baseline:
        # what the reference does for such a step: re-prune the dirty locus from the tips (src/_mcmc_seq.py:746-750)
        sel = sorted(set(touched))[:128]
        cp = CpuPath(cpu_entries_loci([cur[i] for i in sel], [aln[i] for i in sel], taxa), 1)
        try:
            cp.step(1)
            dt, _, cvals = cp.step(3)
        finally:
            cp.close()
        dirty["cpu"] = {"steps_per_s": len(sel) / dt, "kind": cp.kind, "cores": 1,
                        "sample": describe_cpu(cp.kind, 1, "%d of the loci the chain touched (one full re-pruning per step, "
                                               "which is what the reference's scalar cache does)" % len(sel)),
                        "max_rel_err_gpu_vs_cpu": max(abs(eng._felsen[i] - v) / abs(v) for i, v in zip(sel, cvals))}

    # ---- (b) joint chain: both factors on the device ----------------------------------------------------
    eng2 = P.SeqLikelihoodEngine(eng._calcs, species_of, model, context=ctx)
    cur = list(trees)
    heights = [node_heights(t) for t in cur]
    total = eng2.log_likelihood(cur, net, theta)
    n2 = 1000
    lat_locus, lat_theta, accepted, impossible = [], [], 0, 0
    abi_locus = []
    t0 = time.perf_counter()
    done = 0
    for step in range(n2):
        snap = eng2.clone_caches()
        prop_h, loghr = None, 0.0
        if step % 10 == 9:
            new_theta = theta * math.exp(float(rng.uniform(-0.2, 0.2)))
            loghr = math.log(new_theta / theta)
            eng2.invalidate_network()
            prop = cur
            s0 = time.perf_counter()
            new_total = eng2.log_likelihood(prop, net, new_theta)
            lat_theta.append(time.perf_counter() - s0)
        else:
            new_theta = theta
            i = int(rng.integers(0, n_loci))
            if step % 3 == 2:
                o = MV.nni_move(cur[i], heights[i], rng)
                move = None if o is None else (o[0], o[1], 0.0)
            else:
                move = MV.slide_height_move(cur[i], heights[i], rng)
            if move is None:
                continue
            prop = list(cur)
            prop[i] = move[0]
            prop_h = (i, move[1])
            loghr = move[2]
            eng2.invalidate_locus(i)
            s0 = time.perf_counter()
            new_total = eng2.log_likelihood(prop, net, theta)
            lat_locus.append(time.perf_counter() - s0)
            abi_locus.append(ctx.last_timings()["call_ms"])
        done += 1
        impossible += new_total == -math.inf
        if math.log(rng.random()) < (new_total - total) + loghr:
            cur, total, theta = prop, new_total, new_theta
            if prop_h is not None:
                heights[prop_h[0]] = prop_h[1]
            accepted += 1
        else:
            eng2.restore_caches(snap)
    wall = time.perf_counter() - t0
    eng2._commit_pending()
    inc_f, inc_m = list(eng2._felsen), list(eng2._msnc)
    for i in range(n_loci):
        eng2.invalidate_locus(i)
    eng2.invalidate_device_cache()
    full = eng2.log_likelihood(cur, net, theta)
    eng2._commit_pending()
    joint = {
        "what": "both factors of the MCMC_SEQ likelihood on the device through SeqLikelihoodEngine.log_likelihood(gene_trees, "
                "species_net, theta): node-height slide / NNI of one locus, theta scaling every 10th step; Python loop included",
        "loci": n_loci, "species": n_species, "reticulations": 1,
        "logL_evaluations_per_s": done / wall, "accept_rate": accepted / max(done, 1),
        "proposals_without_embedding": int(impossible),
        "median_latency_us_locus_move": float(np.median(lat_locus) * 1e6),
        "median_c_abi_us_locus_move": float(np.median(abi_locus) * 1e3),
        "locus_move_what": "pnb_eval_move: dirty-path pruning and the MSNC density of the moved gene tree queued on two streams, one host synchronisation",
        "median_latency_us_theta_move_all_loci_msnc": float(np.median(lat_theta) * 1e6),
        "cached_equals_recomputed_bitwise": bool(inc_f == list(eng2._felsen) and inc_m == list(eng2._msnc) and full == total),
    }
    if not args.no_cpu_baseline:
        from oracle import msnc_oracle as MO   # checker and CPU baseline only
        netd = net.to_dict()
        k = 40
        t0 = time.perf_counter()
        want = [MO.msnc_log_density(netd, t.parent.tolist(), t.blen.tolist(), t.label, species_of, theta) for t in cur[:k]]
        dt = time.perf_counter() - t0
        joint["cpu_msnc"] = {"gene_trees_per_s": k / dt, "kind": "port", "cores": 1,
                             "sample": "oracle/msnc_oracle.py (Python restatement of the reference DP) on %d gene trees" % k,
                             "max_rel_err_gpu_vs_cpu": max(abs(g - w) / max(1.0, abs(w)) for g, w in zip(eng2._msnc[:k], want))}
    for c in eng._calcs:
        c.close()
    return dirty, joint


# --------------------------------------------------------------------------- reference arm
def run_reference(args, wl, spec, rank):
    """The reference's own CPU implementation of the path on this box's host cores (rank 0 only): the UNMODIFIED
    package from baseline/_ref when present (kind "reference"), else the oracle port.  Each step = one evaluation of
    a bounded sample of the workload; the per-locus objects are built once, outside the clock."""
    if rank != 0:
        return
    from phynetpy_b200 import simulate as S
    cores = os.cpu_count() or 1
    procs = min(cores, 32)
    steps, warm = max(1, min(args.steps, 10)), max(args.warmup, 1)
    if wl == "cfg2":
        from oracle import felsenstein_oracle as O
        n, n_pat = spec["n"], 200_000
        rng = np.random.default_rng(2)
        par, bl = S.random_trees(1, n, rng)
        ft = S.flat_trees(par, bl)[0]
        mat = S.simulate_alignments_counter([ft], int(n_pat * 1.05), PI, RATES, 2, [0])[0]   # the first columns of the GPU arm's data
        first, w, _ = O.compress_fast(mat)
        lut = np.array([O.tip_code_for_char(chr(b)) for b in range(256)], dtype=np.uint8)
        codes = np.ascontiguousarray(lut[mat[:, first[:n_pat]]])
        n_pat = codes.shape[1]
        tip_row = ft.tip_rows({"t%d" % i: i for i in range(n)}).tolist()
        entries = [("codes", codes, w[:n_pat].copy(), tip_row, _tree_tuple(ft))]
        what = "the first %d patterns of the cfg2 workload (1/%d of it)" % (n_pat, max(1, spec["P"] // n_pat))
    else:
        n_loci = min(spec["M"], 512)
        trees, alns = make_locus_data(spec["M"], spec["n"], spec["L"], SEEDS[wl], 0, n_loci, device=False)
        taxa = ["t%d" % i for i in range(spec["n"])]
        entries = cpu_entries_loci(trees, alns, taxa, scales=tuple(1.0 + 0.002 * k for k in range(8)))
        what = "the first %d of the %d loci" % (n_loci, spec["M"])
    cp = CpuPath(entries, procs)
    try:
        for k in range(warm):
            cp.step(1, k)
        dt = 0.0
        ups = 0
        for k in range(steps):
            d, u, _ = cp.step(1, k)
            dt += d
            ups += u
    finally:
        cp.close()
    value = ups / dt
    sample = "each step = one evaluation of " + describe_cpu(cp.kind, cp.procs, what)
    print(json.dumps({
        "impl": "reference", "metric": "site-node partial updates/sec", "value": value, "unit": "updates/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "strong" if wl != "cfg2" else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s" % (wl, spec["desc"]), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "updates/s", "cores": cp.procs, "kind": cp.kind, "sample": sample},
        "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


if __name__ == "__main__":
    main()
