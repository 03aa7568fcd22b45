"""CPU: the host runtime of pnb_eval without a GPU.

``pnb_ctx_create_host`` / ``pnb_plan`` run exactly the host half of an evaluation -- cache
resolution against the committed tree, op-program compilation (register forwarding, stash order,
virtual tip-only nodes, segment numbering), slot assignment, commit / pending / rollback and the
program templates -- and hand the programs back.  A ~60-line NumPy interpreter of the op program
(the semantics of the CUDA kernel, restated) then executes them against persistent per-locus slot
arrays, exactly as the device would, and the result is compared with the oracle.  A compiler bug
(wrong child source, unwritten slot read, stale cache hit) shows up here, on the CPU.
"""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import felsenstein_oracle as O
from phynetpy_b200 import _lib
from phynetpy_b200 import simulate as S
from phynetpy_b200.tree import FlatTree
from conftest import load_golden

CASES = load_golden("felsen_cases.json")["cases"]
K_TIP, K_MEM, K_REG, K_STK = 0, 1, 2, 3
F_FIRST, F_LAST, F_STORE, F_PUSH, F_HOLD = 0x4, 0x8, 0x10, 0x20, 0x40
SEG_MAX_OPS = 192
BITS = np.array([[(c >> j) & 1 for j in range(4)] for c in range(16)], dtype=np.float64)


class Planner:
    """Host-only context + loci + the simulated device state (slot arrays) of each locus."""

    def __init__(self):
        self.lib = _lib.load()
        self.ctx = C.c_void_p()
        _lib.check(self.lib.pnb_ctx_create_host(C.byref(self.ctx)))
        self.loci = []      # (handle, OracleLocus, codes, {slot: partial})
        self.models = {}

    def close(self):
        for h, *_ in self.loci:
            self.lib.pnb_locus_destroy(h)
        for h, _ in self.models.values():
            self.lib.pnb_model_destroy(h)
        self.lib.pnb_ctx_destroy(self.ctx)

    def add_locus(self, alignment):
        ol = O.OracleLocus(dict(alignment))
        codes = np.ascontiguousarray(ol.codes())
        h = C.c_void_p()
        w = np.ascontiguousarray(ol.int_weights)
        _lib.check(self.lib.pnb_locus_create(self.ctx, codes.shape[0], codes.shape[1], _lib.ptr(codes),
                                             max(codes.shape[1], 1), _lib.ptr(w), C.byref(h)))
        self.loci.append((h, ol, codes, {}))
        return len(self.loci) - 1

    def model(self, om, key):
        if key not in self.models and om.jc_closed_form:
            h = C.c_void_p()
            _lib.check(self.lib.pnb_model_create_jc69(self.ctx, C.byref(h)))
            self.models[key] = (h, om)
        if key not in self.models:
            h = C.c_void_p()
            _lib.check(self.lib.pnb_model_create_eigen(
                self.ctx, _lib.ptr(np.ascontiguousarray(om._evals)), _lib.ptr(np.ascontiguousarray(om._U)),
                _lib.ptr(np.ascontiguousarray(om._Uinv)), _lib.ptr(np.ascontiguousarray(om.pi)), C.byref(h)))
            self.models[key] = (h, om)
        return self.models[key]

    def plan(self, jobs, model_key, flags=0, site_rate=1.0):
        """jobs: [(locus index, FlatTree)] -> per-job (info, ops, t)."""
        mh, _ = self.models[model_key]
        handles = np.array([self.loci[i][0].value for i, _ in jobs], dtype=np.uint64)
        sizes = [t.n_nodes for _, t in jobs]
        node_off = np.zeros(len(jobs) + 1, dtype=np.int64)
        node_off[1:] = np.cumsum(sizes)
        parent = np.ascontiguousarray(np.concatenate([t.parent for _, t in jobs]), dtype=np.int32)
        blen = np.ascontiguousarray(np.concatenate([t.blen for _, t in jobs]), dtype=np.float64)
        rows = []
        for i, t in jobs:
            row_of = {name: r for r, name in enumerate(self.loci[i][1].taxa)}
            rows.append(t.tip_rows(row_of))
        tip_row = np.ascontiguousarray(np.concatenate(rows), dtype=np.int32)
        total = int(node_off[-1])
        info = np.zeros((len(jobs), 8), dtype=np.int32)
        ops = np.zeros((total, 4), dtype=np.int32)
        t_out = np.zeros(total, dtype=np.float64)
        _lib.check(self.lib.pnb_plan(self.ctx, mh, len(jobs), _lib.ptr(handles), _lib.ptr(node_off), _lib.ptr(parent),
                                     _lib.ptr(blen), _lib.ptr(tip_row), float(site_rate), flags,
                                     _lib.ptr(info), _lib.ptr(ops), _lib.ptr(t_out)))
        out = []
        for j in range(len(jobs)):
            o = int(node_off[j]) - j
            n = int(info[j, 0])
            out.append((info[j].copy(), ops[o:o + n].copy(), t_out[o:o + n].copy()))
        return out

    def commit(self, idx):
        h = np.array([self.loci[i][0].value for i in idx], dtype=np.uint64)
        _lib.check(self.lib.pnb_commit(self.ctx, len(idx), _lib.ptr(h)))

    def rollback(self, idx):
        h = np.array([self.loci[i][0].value for i in idx], dtype=np.uint64)
        _lib.check(self.lib.pnb_rollback(self.ctx, len(idx), _lib.ptr(h)))

    def run(self, locus_idx, planned, model_key):
        """Interpret one job's program on the simulated device state; returns logL."""
        info, ops, t = planned
        _, ol, codes, slots = self.loci[locus_idx]
        om = self.models[model_key][1]
        n_ops, n_tip_ops, root_kind, root_src, stk_need, n_dirty, rows_cap, cached = (int(x) for x in info)
        P = codes.shape[1]
        acc = None
        complete = False          # acc holds a finished node (what a REG op may consume)
        stack = {}
        hold = None               # PNB_HOLD builds: the ONE register-held sibling (K_STK with src < 0 reads it)
        seen_aux = set()
        for k in range(n_ops):
            flags, src, dst, aux = (int(x) for x in ops[k])
            flags &= 0xFFFFFFFF
            kind = flags & 3
            if k % SEG_MAX_OPS == 0:
                seen_aux = set()
            Pm = om.p_matrix(float(t[k]))
            if kind == K_TIP:
                if src >= 0:
                    assert 0 <= aux < max(rows_cap, 1) and aux not in seen_aux, "staging row clash"
                    seen_aux.add(aux)
                    child = BITS[codes[src]]
                else:
                    child = np.ones((P, 4))
            elif kind == K_MEM:
                assert src in slots, "op %d reads slot %d that was never written" % (k, src)
                child = slots[src]
            elif kind == K_REG:
                assert flags & F_FIRST and complete, "REG must be the first op after a completed node"
                child = acc
            elif src < 0:
                assert hold is not None, "HOLD registers read while empty"
                child, hold = hold, None
            else:
                assert src in stack, "stash %d read before it was pushed" % src
                assert src < stk_need
                child = stack[src]
            v = child @ Pm.T
            if flags & F_FIRST:
                acc = v
            else:
                assert not complete or kind != K_REG
                acc = acc * v
            complete = bool(flags & F_LAST)
            if flags & F_LAST:
                if flags & F_STORE:
                    slots[dst] = acc.copy()
                if flags & F_HOLD:
                    assert hold is None, "a second node wants the HOLD registers while they are in use"
                    hold = acc.copy()
                if flags & F_PUSH:
                    assert ((flags >> 8) & 0xFF) < stk_need, "stash index beyond the capacity the job announced"
                    stack[(flags >> 8) & 0xFF] = acc.copy()
        if root_kind == K_REG:
            assert n_ops > 0 and complete
            root = acc
        elif root_kind == K_MEM:
            assert root_src in slots
            root = slots[root_src]
        else:
            root = BITS[codes[root_src]] if root_src >= 0 else np.ones((P, 4))
        site = np.maximum(root @ om.pi, 1e-300)
        return float(np.dot(np.log(site), ol.weights))


VIRT_TIPS = [2, 4, 7]      # PNB_VIRT_TIPS: subtrees of at most this many tips are recomputed, never stored


@pytest.fixture(params=VIRT_TIPS, ids=["virt%d" % v for v in VIRT_TIPS])
def planner(request, monkeypatch):
    monkeypatch.setenv("PNB_VIRT_TIPS", str(request.param))   # read when the planning context is created
    p = Planner()
    p.virt_tips = request.param
    yield p
    p.close()


def _tips_below(ft):
    nc = ft.n_children()
    tips = (nc == 0).astype(np.int64)
    order = []
    stack = [int(np.nonzero(ft.parent < 0)[0][0])]
    kids = [[] for _ in range(ft.n_nodes)]
    for i, p in enumerate(ft.parent.tolist()):
        if p >= 0:
            kids[p].append(i)
    while stack:
        v = stack.pop()
        order.append(v)
        stack.extend(kids[v])
    for v in reversed(order):
        for c in kids[v]:
            tips[v] += tips[c]
    return tips


def _omodel(spec):
    if spec["kind"] == "JC69":
        return O.jc69()
    if spec["kind"] == "HKY85":
        return O.hky85(spec["kappa"], spec["pi"])
    return O.gtr(spec["pi"], spec["rates"])


def _flat(t):
    return FlatTree(t["parent"], t["blen"], t["label"])


def _otree(ft):
    return O.OracleTree(ft.parent.tolist(), [None if math.isnan(b) else float(b) for b in ft.blen], ft.label)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_compiled_program_reproduces_reference_logL(planner, case):
    """Every golden case: compile (store mode), interpret, compare with the reference's logL; then the
    same tree again must be a pure cache hit, and NOSTORE must give the same number."""
    om = _omodel(case["model"])
    planner.model(om, "m")
    i = planner.add_locus(case["alignment"])
    ft = _flat(case["tree"])
    (p,) = planner.plan([(i, ft)], "m", 0, case["site_rate"])
    n_internal = int((ft.n_children() > 0).sum())
    assert p[0][5] == n_internal
    got = planner.run(i, p, "m")
    assert got == pytest.approx(case["logL"], rel=1e-11, abs=1e-11)
    (p2,) = planner.plan([(i, ft)], "m", 0, case["site_rate"])
    assert p2[0][5] == 0                             # nothing changed
    nc = ft.n_children()
    root = int(np.nonzero(ft.parent < 0)[0][0])
    root_is_tip_only = nc[root] > 0 and all(nc[c] == 0 for c in np.nonzero(ft.parent == root)[0])
    root_is_virtual = nc[root] > 0 and (root_is_tip_only or _tips_below(ft)[root] <= planner.virt_tips)
    if root_is_virtual:
        assert p2[0][2] == K_REG and p2[0][0] == ft.n_nodes - 1   # a small-subtree root is recomputed from the codes
    elif n_internal:
        assert p2[0][2] == K_MEM and p2[0][0] == 0          # read back from its slot, no ops
    assert planner.run(i, p2, "m") == got
    (p3,) = planner.plan([(i, ft)], "m", _lib.EVAL_NOSTORE | _lib.EVAL_FULL, case["site_rate"])
    assert planner.run(i, p3, "m") == pytest.approx(case["logL"], rel=1e-11, abs=1e-11)
    assert not any(int(f) & F_STORE for f in p3[1][:, 0])


def test_dirty_path_sequences_commit_and_rollback(planner):
    rng = np.random.default_rng(8)
    n, L = 14, 120
    par, bl = S.random_trees(1, n, rng, height=0.5)
    aln = S.simulate_alignments(par, bl, L, S.DEFAULT_PI, S.DEFAULT_RATES, rng)[0]
    d = S.to_alignment_dict(aln)
    om = O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES)
    planner.model(om, "m")
    i = planner.add_locus(list(d.items()))
    ol = planner.loci[i][1]
    ft = S.flat_trees(par, bl)[0]
    (p,) = planner.plan([(i, ft)], "m")
    assert planner.run(i, p, "m") == pytest.approx(O.log_likelihood(ol, _otree(ft), om), rel=1e-12)
    cur = ft
    for step in range(25):
        node = int(rng.integers(0, 2 * n - 2))
        b = cur.blen.copy()
        b[node] *= float(rng.uniform(0.5, 2.0))
        prop = cur.with_lengths(b)
        depth, q = 0, cur.parent[node]
        while q >= 0:
            depth, q = depth + 1, cur.parent[q]
        (p,) = planner.plan([(i, prop)], "m", _lib.EVAL_PENDING)
        assert p[0][5] == depth                               # exactly the path to the root
        got = planner.run(i, p, "m")
        assert got == pytest.approx(O.log_likelihood(ol, _otree(prop), om), rel=1e-12)
        with pytest.raises(_lib.PhyNetB200Error, match="pending"):
            planner.plan([(i, prop)], "m", _lib.EVAL_PENDING)
        if step % 3 == 0:
            planner.rollback([i])                             # reject: the old tree is still fully cached
            (p0,) = planner.plan([(i, cur)], "m", _lib.EVAL_PENDING)
            assert p0[0][5] == 0
            planner.rollback([i])
        else:
            planner.commit([i])
            cur = prop
    # an NNI re-hangs two subtrees: only the two touched nodes and their ancestors are recomputed
    parent = cur.parent.copy()
    v = n
    while (cur.parent == v).sum() == 0 or cur.parent[v] < 0:
        v += 1
    u = int(parent[v])
    child = int(np.nonzero(parent == v)[0][0])
    sib = [int(c) for c in np.nonzero(parent == u)[0] if c != v][0]
    parent[child], parent[sib] = u, v
    nni = FlatTree(parent, cur.blen, cur.label)
    (p,) = planner.plan([(i, nni)], "m")
    assert 0 < p[0][5] < n - 1
    assert planner.run(i, p, "m") == pytest.approx(O.log_likelihood(ol, _otree(nni), om), rel=1e-12)


def test_templates_candidates_and_mode_switches(planner):
    rng = np.random.default_rng(21)
    n = 11
    par, bl = S.random_trees(2, n, rng, height=0.4)
    aln = S.simulate_alignments(par[:1], bl[:1], 90, S.DEFAULT_PI, S.DEFAULT_RATES, rng)[0]
    om = O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES)
    planner.model(om, "m")
    om2 = O.hky85(3.0, [0.2, 0.3, 0.3, 0.2])
    planner.model(om2, "m2")
    i = planner.add_locus(list(S.to_alignment_dict(aln).items()))
    ol = planner.loci[i][1]
    t0, t1 = S.flat_trees(par, bl)
    first = None
    for scale in (1.0, 0.5, 2.0, 4.0):                         # same topology, new lengths: template path
        t = t0.with_lengths(t0.blen * scale)
        (p,) = planner.plan([(i, t)], "m", _lib.EVAL_FULL)
        assert p[0][5] == n - 1
        if first is None:
            first = p[1].copy()
        assert np.array_equal(p[1], first)                     # the program is the topology's; only t changes
        assert planner.run(i, p, "m") == pytest.approx(O.log_likelihood(ol, _otree(t), om), rel=1e-12)
    # the cache image restored from the template belongs to the LAST tree evaluated
    b = t0.blen * 4.0
    b[2] *= 1.1
    (p,) = planner.plan([(i, t0.with_lengths(b))], "m")
    assert 0 < p[0][5] < n - 1
    assert planner.run(i, p, "m") == pytest.approx(O.log_likelihood(ol, _otree(t0.with_lengths(b)), om), rel=1e-12)
    # candidates for one locus in one NOSTORE batch: two topologies x scales, the locus repeated
    cands = [(i, t.with_lengths(t.blen * s)) for t in (t0, t1) for s in (0.25, 1.0, 4.0)]
    plans = planner.plan(cands, "m", _lib.EVAL_NOSTORE)
    for (_, t), p in zip(cands, plans):
        assert planner.run(i, p, "m") == pytest.approx(O.log_likelihood(ol, _otree(t), om), rel=1e-12)
        assert not any(int(f) & F_STORE for f in p[1][:, 0])
    with pytest.raises(ValueError, match="twice"):
        planner.plan([(i, t0), (i, t1)], "m")                  # storing batches hold a locus once
    # another model, or the other numerical mode, cannot reuse the cached partials
    (p,) = planner.plan([(i, t0.with_lengths(b))], "m2")
    assert p[0][5] == n - 1
    assert planner.run(i, p, "m2") == pytest.approx(O.log_likelihood(ol, _otree(t0.with_lengths(b)), om2), rel=1e-12)
    (p,) = planner.plan([(i, t0.with_lengths(b))], "m2", _lib.EVAL_RESCALE)
    assert p[0][5] == n - 1


def test_long_programs_are_segmented_and_large_batches(planner):
    rng = np.random.default_rng(4)
    n = 150                                                    # 298 ops -> 2 segments of <= 192
    par, bl = S.random_trees(1, n, rng, height=0.3)
    aln = S.simulate_alignments(par, bl, 40, S.DEFAULT_PI, S.DEFAULT_RATES, rng)[0]
    om = O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES)
    planner.model(om, "m")
    i = planner.add_locus(list(S.to_alignment_dict(aln).items()))
    ft = S.flat_trees(par, bl)[0]
    for flags in (0, _lib.EVAL_NOSTORE | _lib.EVAL_FULL):
        (p,) = planner.plan([(i, ft)], "m", flags)
        assert p[0][0] == 2 * (n - 1) and p[0][6] <= SEG_MAX_OPS
        assert planner.run(i, p, "m") == pytest.approx(O.log_likelihood(planner.loci[i][1], _otree(ft), om), rel=1e-12)
    # many loci in one batch, ragged sizes
    jobs, exp = [], []
    for k in range(40):
        nk = int(rng.integers(3, 20))
        pk, bk = S.random_trees(1, nk, rng, height=0.3)
        ak = S.simulate_alignments(pk, bk, int(rng.integers(5, 60)), S.DEFAULT_PI, S.DEFAULT_RATES, rng)[0]
        idx = planner.add_locus(list(S.to_alignment_dict(ak).items()))
        t = S.flat_trees(pk, bk)[0]
        jobs.append((idx, t))
        exp.append(O.log_likelihood(planner.loci[idx][1], _otree(t), om))
    for (idx, _), p, e in zip(jobs, planner.plan(jobs, "m"), exp):
        assert planner.run(idx, p, "m") == pytest.approx(e, rel=1e-12)


def test_invalid_trees_are_rejected_with_value_errors(planner):
    om = O.gtr(S.DEFAULT_PI, S.DEFAULT_RATES)
    planner.model(om, "m")
    i = planner.add_locus([("A", "ACGT"), ("B", "ACGA")])
    lab = ["A", "B", None]
    with pytest.raises(ValueError, match="more than one root"):
        planner.plan([(i, FlatTree([-1, -1, 0], [0.1] * 3, lab))], "m")
    with pytest.raises(ValueError, match="no root"):
        planner.plan([(i, FlatTree([1, 0, 0], [0.1] * 3, lab))], "m")
    with pytest.raises(ValueError, match="not connected"):
        planner.plan([(i, FlatTree([2, 2, -1, 4, 3], [0.1] * 5, lab + [None, None]))], "m")
    with pytest.raises(ValueError, match="out of range"):
        planner.plan([(i, FlatTree([2, 7, -1], [0.1] * 3, lab))], "m")
    good = FlatTree([2, 2, -1], [0.1, 0.2, math.nan], lab)
    (p,) = planner.plan([(i, good)], "m")                       # the failures left no pending state behind
    assert planner.run(i, p, "m") == pytest.approx(
        O.log_likelihood(planner.loci[i][1], _otree(good), om), rel=1e-12)
