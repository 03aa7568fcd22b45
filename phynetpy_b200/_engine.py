"""Batched / dirty-path likelihood engine: the drop-in for the Felsenstein half of
the reference's ``_SeqLikelihoodEngine`` (``src/_mcmc_seq.py:647-772``).

Same seam as the reference: ``log_likelihood(gene_trees, species_net, theta)``,
``invalidate_locus(i)``, ``invalidate_network()``, ``clone_caches()`` /
``restore_caches(snapshot)`` and the scalar per-locus caches ``_felsen`` /
``_msnc`` -- so every MCMC_SEQ operator and its undo closure work unmodified.
What changes is how the ``None`` entries of ``_felsen`` are filled: all dirty loci
go to the GPU in ONE ``pnb_eval`` call, and inside each locus only the nodes whose
subtree changed are recomputed (per-node partials stay resident in HBM).

The MSNC coalescent density (the other factor, ``src/_msnc_density.py``) is filled
the same way when a species network is passed: every locus whose ``_msnc`` entry is
``None`` goes to kernel K-m in ONE ``pnb_msnc_eval`` call (``_msnc_density.py`` here).
A host callable can still be plugged in as ``msnc_fn`` instead.

Multi-GPU (one process per GPU): loci are sharded in contiguous ranges; the pruning
kernel writes a rank's per-locus values straight into its slice of a device vector
(``PNB_EVAL_OUT_DEVICE``) and ONE in-place ``ncclAllGather`` on the same stream gives
every rank all values (``ShardPlan.allgather_slices_``); one pinned D2H copy brings the
vector to the host, which sums it in locus order, so the total does not depend on the
number of GPUs.  When every local locus is dirty and the topologies are the ones of the
previous such evaluation, the loci go through a resident batch (``pnb_batch_*``): only
branch lengths travel.
"""
from __future__ import annotations

import math
from typing import Any, Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._msnc_density import SpeciesNetwork, as_species_network, msnc_eval_packed, _check_theta
from ._seq_likelihood import FelsensteinCalculator, _device_kind, _explicit_P_edges, _model_handle
from .device import DeviceContext, ResidentBatch, default_context, pack_trees
from .distributed import ShardPlan
from .tree import FlatTree, flatten

MsncFn = Callable[[int, Any, Any, float], float]


class _SummedList(list):
    """A per-locus cache list (``_felsen`` / ``_msnc``) that knows its sum.  The reference returns the total as a loop over all
    loci on every evaluation (``src/_mcmc_seq.py:745-761``); after a single-locus move that loop -- even as two C-level
    ``sum`` calls -- costs more than everything else the host does for the step (20 of 92 us at 1,000 loci).  Here the sum is
    kept per block of 32 entries: an item assignment marks its block, ``total()`` re-adds the marked blocks and then the
    block sums, always in the same order -- so the total is a function of the CONTENT alone (a cached state and a
    from-scratch evaluation of the same state give the same bits), and a ``None`` written by anybody, through any
    reference to the list, still raises ``TypeError`` in ``total()`` exactly as ``sum`` would."""

    B = 32

    def __init__(self, values=()):
        super().__init__(values)
        self._blocks = [None] * ((len(self) + self.B - 1) // self.B)

    def _all_dirty(self):
        self._blocks = [None] * ((len(self) + self.B - 1) // self.B)

    def __setitem__(self, i, v):
        if type(i) is int:
            list.__setitem__(self, i, v)
            self._blocks[(i if i >= 0 else i + len(self)) // self.B] = None
            return
        n = len(self)
        list.__setitem__(self, i, v)
        if isinstance(i, slice) and len(self) == n:     # a run of entries rewritten in place: only its blocks
            start, stop, step = i.indices(n)
            if step == 1 and stop > start:
                for b in range(start // self.B, (stop - 1) // self.B + 1):
                    self._blocks[b] = None
                return
            if stop <= start and step == 1:
                return
        elif not isinstance(i, slice):                  # an integer-like index (numpy integer)
            k = int(i)
            self._blocks[(k if k >= 0 else k + n) // self.B] = None
            return
        self._all_dirty()

    def _mutator(name):   # every other way of changing a list: correct, not fast
        def f(self, *a, **k):
            r = getattr(list, name)(self, *a, **k)
            self._all_dirty()
            return r
        f.__name__ = name
        return f

    for _n in ("__delitem__", "__iadd__", "__imul__", "append", "extend", "insert", "pop", "remove", "clear", "sort", "reverse"):
        locals()[_n] = _mutator(_n)
    del _n, _mutator

    def copy(self):
        c = _SummedList.__new__(_SummedList)
        list.extend(c, self)                 # C-level copy; the block sums come along
        c._blocks = self._blocks.copy()
        return c

    def __reduce__(self):   # pickled (run_parallel_chains) as its content
        return (_SummedList, (list(self),))

    def total(self):
        if len(self) <= 2 * self.B:
            return sum(self)
        bl, B = self._blocks, self.B
        for b, v in enumerate(bl):
            if v is None:
                bl[b] = sum(self[b * B:(b + 1) * B])    # TypeError when an entry is None, as sum(list) would
        return sum(bl)


def _list_total(values) -> float:
    return values.total() if type(values) is _SummedList else sum(values)


def _copy_cache(values) -> list:
    return values.copy() if type(values) is _SummedList else _SummedList(values)


class SeqLikelihoodEngine:
    """Incremental ``sum_i log P(S_i | g_i) [+ log P(g_i | Psi)]`` over all loci.

    Attributes:
        n_loci: number of loci (global, across all ranks).
    """

    def __init__(self, loci: Sequence[Any], species_of: Optional[dict] = None, model: Any = None,
                 branch_thetas: Optional[dict] = None, *, msnc_fn: Optional[MsncFn] = None,
                 context: Optional[DeviceContext] = None, shard: Optional[ShardPlan] = None,
                 rescale: bool = False, resident_batches: bool = False) -> None:
        """``loci``: per-locus alignments (``dict`` label -> sequence) or ready
        :class:`FelsensteinCalculator` objects.  With ``shard`` only this rank's loci
        are made resident (entries outside the shard may be ``None``).

        ``resident_batches``: evaluations in which EVERY local locus is dirty go through a resident batch
        (``pnb_batch_*``: op programs stay in HBM, only branch lengths travel, every node is recomputed and the
        result is committed at once).  The right choice when such evaluations change all branch lengths (height
        rescalings, model changes, throughput runs); leave it off when they change a branch or two per locus --
        the default path recomputes only the changed paths and keeps the evaluation pending for a rollback."""
        self.n_loci = len(loci)
        self._species_of = species_of
        self._model = model
        self._branch_thetas = branch_thetas
        self._msnc_fn = msnc_fn
        self._ctx = context
        self._shard = shard or ShardPlan.single(self.n_loci)
        lo, hi = self._shard.local_range()
        self._calcs: List[Optional[FelsensteinCalculator]] = [None] * self.n_loci
        for i in range(lo, hi):
            a = loci[i]
            self._calcs[i] = a if isinstance(a, FelsensteinCalculator) else FelsensteinCalculator(a, context=context)
        self._felsen: List[Optional[float]] = _SummedList([None] * self.n_loci)
        self._msnc: List[Optional[float]] = _SummedList([None] * self.n_loci)
        self._pending: List[int] = []   # loci with an uncommitted device evaluation
        self._dirty: set = set(range(lo, hi))  # local loci whose _felsen entry is None
        self._dirty_any: set = set(range(self.n_loci))   # loci whose entry is None on ANY rank (sharded engines)
        self._snap_dirty: dict = {}            # id(snapshot list) -> dirty sets at clone time
        self._msnc_dirty: set = set()          # loci whose _msnc entry is None (when not all of them are)
        self._msnc_all = True                  # every _msnc entry is None
        self._full_next = False
        self._mode = _lib.EVAL_RESCALE if rescale else 0  # numerical mode (beyond the reference when set)
        self.resident_batches = bool(resident_batches)
        self._stats: Optional[dict] = {}
        self._stats_ctx: Any = None
        # MSNC side: the network view is rebuilt when a different network OBJECT arrives (the reference
        # keys on identity too, src/_mcmc_seq.py:705-713); per-locus species ids are keyed by the label list
        self._net_obj: Any = None
        self._net: Optional[SpeciesNetwork] = None
        self._sp_ids: List[Optional[Tuple[Any, np.ndarray]]] = [None] * self.n_loci
        self._store: Optional[dict] = None     # packed gene trees of all loci (see _packed_all)
        self._batch: Optional[ResidentBatch] = None   # resident programs of the last all-local-loci evaluation
        self._batch_key: Optional[tuple] = None
        self._sh: Optional[dict] = None        # sharded mode: device vector, pinned host copy, stream (see _shard_state)
        self._scan_pack: Optional[dict] = None   # gene trees of all loci at every height scale, packed (network_only_msnc_score)

    @property
    def last_stats(self) -> dict:
        """Counters of the device call behind the engine's last evaluation (``DeviceContext.last_stats``), read from the
        library when asked for: a chain step does not pay a second ABI call for numbers nobody looks at."""
        if self._stats is None:
            self._stats = self._stats_ctx.last_stats() if self._stats_ctx is not None else {}
        return self._stats

    @last_stats.setter
    def last_stats(self, value: dict) -> None:
        self._stats = value

    # -- reference seam (src/_mcmc_seq.py:696-703, :766-772) ------------------
    def invalidate_locus(self, i: int) -> None:
        self._felsen[i] = None
        self._msnc[i] = None
        lo, hi = self._shard.local_range()
        if lo <= i < hi:
            self._dirty.add(i)
        self._dirty_any.add(i)
        self._msnc_dirty.add(i)

    def invalidate_network(self) -> None:
        self._msnc = _SummedList([None] * self.n_loci)
        self._msnc_all = True
        self._msnc_dirty.clear()

    def clone_caches(self) -> Tuple[list, list]:
        # (copies that carry their block sums: a restored snapshot does not pay the full sum again)
        snap = (_copy_cache(self._felsen), _copy_cache(self._msnc))
        if len(self._snap_dirty) > 8:
            self._snap_dirty.clear()
        self._snap_dirty[id(snap[0])] = (snap[0], frozenset(self._dirty), frozenset(self._msnc_dirty), self._msnc_all,
                                         frozenset(self._dirty_any))
        return snap

    def restore_caches(self, snapshot: Tuple[list, list]) -> None:
        """Reject path: scalar caches come back from the snapshot and the device
        discards the partials written by the rejected proposal."""
        self._felsen, self._msnc = _copy_cache(snapshot[0]), _copy_cache(snapshot[1])
        known = self._snap_dirty.get(id(snapshot[0]))
        if known is not None and known[0] is snapshot[0]:
            self._dirty = set(known[1])
            self._msnc_dirty, self._msnc_all = set(known[2]), known[3]
            self._dirty_any = set(known[4])
        else:  # a snapshot we did not hand out: recount
            lo, hi = self._shard.local_range()
            self._dirty = {i for i in range(lo, hi) if self._felsen[i] is None}
            self._dirty_any = {i for i, v in enumerate(self._felsen) if v is None}
            self._recount_msnc()
        self._rollback_pending()

    def invalidate_device_cache(self) -> None:
        """Force the next evaluation to recompute every node of every dirty locus."""
        self._full_next = True

    # -- device bookkeeping ---------------------------------------------------
    def _context(self) -> DeviceContext:
        if self._ctx is None:
            self._ctx = default_context()
        return self._ctx

    def _handles(self, idx: Sequence[int]) -> np.ndarray:
        out = np.empty(len(idx), dtype=np.uint64)
        for k, i in enumerate(idx):
            out[k] = self._calcs[i]._locus_handle().value
        return out

    def _commit_pending(self) -> None:
        if self._pending:
            self._context().commit(self._handles(self._pending))
            self._pending = []

    def _rollback_pending(self) -> None:
        if self._pending:
            self._context().rollback(self._handles(self._pending))
            self._pending = []

    # -- evaluation -----------------------------------------------------------
    def felsenstein_batch(self, idx: Sequence[int], gene_trees: Sequence[Any], *, flags: int = 0,
                          site_rate: float = 1.0, out_device_ptr: Optional[int] = None) -> Optional[np.ndarray]:
        """log P(S_i | g_i) for the given local loci in one device call."""
        ctx = self._context()
        flats = [flatten(gene_trees[i]) for i in idx]
        node_off, parent, blen, tip_row = pack_trees(flats, [self._calcs[i]._row_of for i in idx])
        P_edge = None
        if _device_kind(self._model) == "explicit":
            P_edge = _explicit_P_edges(self._model, blen, site_rate)
        extra = {} if out_device_ptr is None else {"out_device_ptr": out_device_ptr}
        out = ctx.eval(_model_handle(self._model, ctx), self._handles(idx), node_off, parent, blen, tip_row,
                       site_rate, flags, P_edge, **extra)
        self._stats, self._stats_ctx = None, ctx      # fetched when somebody reads last_stats
        return out

    def log_likelihood(self, gene_trees: Sequence[Any], species_net: Any = None, theta: float = 0.0,
                       *, rescan: bool = False) -> float:
        """Total log-likelihood; ``-inf`` when not finite (reference :723-764)."""
        # whatever was pending belongs to an accepted proposal: it becomes the committed state
        self._commit_pending()
        lo, hi = self._shard.local_range()
        if rescan:  # entries were set to None behind our back (a caller sharing the list): recount
            self._dirty = {i for i in range(lo, hi) if self._felsen[i] is None}
            self._dirty_any = {i for i, v in enumerate(self._felsen) if v is None}
            self._recount_msnc()
        dirty = sorted(i for i in self._dirty if self._felsen[i] is None)
        self._dirty.clear()
        flags = _lib.EVAL_PENDING | self._mode | (_lib.EVAL_FULL if self._full_next else 0)
        self._full_next = False
        if (self._shard.world_size == 1 and len(dirty) == 1 and species_net is not None and self._msnc_fn is None
                and not self._msnc_all and self._msnc_dirty == {dirty[0]} and self._msnc[dirty[0]] is None
                and _device_kind(self._model) != "explicit" and hasattr(self._context(), "eval_move")):
            self._eval_move(dirty[0], gene_trees[dirty[0]], species_net, theta, flags)   # a gene-tree move: one device call
        elif self._shard.world_size == 1:
            if dirty:
                if self.resident_batches and len(dirty) == hi - lo and len(dirty) >= self.BATCH_MIN_LOCI \
                        and _device_kind(self._model) != "explicit" and not (flags & _lib.EVAL_FULL):
                    vals = self._eval_all_local(dirty, gene_trees)        # resident batch: committed at once
                else:
                    vals = self.felsenstein_batch(dirty, gene_trees, flags=flags)
                    self._pending = list(dirty)
                if len(dirty) == dirty[-1] - dirty[0] + 1:
                    self._felsen[dirty[0]:dirty[-1] + 1] = vals.tolist()     # a contiguous run: one slice assignment
                else:
                    for i, v in zip(dirty, vals.tolist()):
                        self._felsen[i] = v
        else:
            self._log_likelihood_sharded(dirty, gene_trees, flags)
        try:
            return self._total(gene_trees, species_net, theta)
        except TypeError:
            # an entry is still None: it was invalidated without invalidate_locus() (a caller that shares
            # the list); recount the dirty set once and evaluate again
            if rescan:
                raise
            return self.log_likelihood(gene_trees, species_net, theta, rescan=True)

    def _eval_move(self, i: int, gene_tree: Any, species_net: Any, theta: float, flags: int) -> None:
        """A gene-tree move dirtied locus ``i`` alone: its Felsenstein factor (dirty-path pruning) and its MSNC
        density in ONE device call with one host synchronisation (``pnb_eval_move``); the reference computes them
        one after the other in its loop over loci (``src/_mcmc_seq.py:746-760``)."""
        theta = _check_theta(theta)
        ctx = self._context()
        net = self._network(species_net)
        ft = flatten(gene_tree)
        calc = self._calcs[i]
        f, m = ctx.eval_move(_model_handle(self._model, ctx), calc._locus_handle(), ft.parent, ft.blen, ft.tip_rows(calc._row_of),
                             1.0, flags, net.handle(ctx, self._branch_thetas), self._species_ids(i, ft, net), theta)
        self._felsen[i] = f
        self._msnc[i] = m
        self._pending = [i]
        self._msnc_dirty.clear()
        self._stats, self._stats_ctx = None, ctx      # fetched when somebody reads last_stats

    BATCH_MIN_LOCI = 32   # all-local-loci evaluations of at least this many loci go through a resident batch

    def _eval_all_local(self, idx: Sequence[int], gene_trees: Sequence[Any], out_device_ptr: Optional[int] = None,
                        peer_ptrs: Sequence[int] = (), barrier: Optional[tuple] = None) -> Optional[np.ndarray]:
        """Every local locus at once through a resident batch: the op programs of the previous such evaluation
        are reused when the topologies are the same (one memcmp of the packed parent arrays), so only branch
        lengths travel.  The evaluation is committed at once: the loci's partial caches are content-addressed,
        so a proposal that is rejected afterwards costs the next evaluation of a locus a full recomputation of
        that locus, never a wrong value."""
        ctx = self._context()
        flats = [flatten(gene_trees[i]) for i in idx]
        node_off, parent, blen, tip_row = pack_trees(flats, [self._calcs[i]._row_of for i in idx])
        mh = _model_handle(self._model, ctx)
        key = self._batch_key
        if self._batch is None or key is None or key[0] != self._mode or key[1].shape != parent.shape or \
                not np.array_equal(key[1], parent) or not np.array_equal(key[2], tip_row) or not np.array_equal(key[3], node_off):
            if self._batch is not None:
                self._batch.close()
            self._batch = ResidentBatch(ctx, mh, self._handles(idx), node_off, parent, tip_row, self._mode)
            self._batch_key = (self._mode, parent, tip_row, node_off)
        try:
            self._batch.set_peer_outputs(list(peer_ptrs))
            if barrier is not None:
                self._batch.set_peer_barrier(*barrier)
            out = self._batch.eval(mh, blen, 1.0, out_device_ptr)
        except _lib.PhyNetB200Error:
            # a locus was re-allocated or destroyed underneath the batch: build it again, once
            self._batch.close()
            self._batch = ResidentBatch(ctx, mh, self._handles(idx), node_off, parent, tip_row, self._mode)
            self._batch.set_peer_outputs(list(peer_ptrs))
            if barrier is not None:
                self._batch.set_peer_barrier(*barrier)
            out = self._batch.eval(mh, blen, 1.0, out_device_ptr)
        self._stats, self._stats_ctx = None, ctx      # fetched when somebody reads last_stats
        return out

    # -- sharded evaluation: device vector + in-place all-gather ------------------------------------------
    def _shard_state(self) -> dict:
        if self._sh is None:
            import torch
            plan = self._shard
            per = plan.slice_len
            on_gpu = torch.cuda.is_available() and getattr(self._context(), "device", None) is not None \
                and isinstance(self._context(), DeviceContext)
            sh = {"per": per, "index": plan.gathered_index(), "gpu": on_gpu}
            if on_gpu:
                dev = torch.device("cuda", self._context().device)
                sh["stream"] = torch.cuda.Stream(device=dev)
                self._context().set_stream(sh["stream"].cuda_stream)   # kernels and the collective share one stream
                # preferred: the vector in symmetric memory, so that the kernel of a full evaluation stores every locus'
                # value straight into all peers' copies (fused collective); plain memory + NCCL all-gather otherwise
                with torch.cuda.device(dev):
                    sym = plan.symmetric_vector()
                n = per * plan.world_size
                if sym is not None:
                    # two copies of the vector, used alternately by consecutive evaluations (ShardPlan.symmetric_vector)
                    sh["vecs"], sh["sym"] = [sym[0][:n], sym[0][n:2 * n]], sym[1]
                    sh["peers"] = [[sym[1].buffer_ptrs[r] + 8 * (c * n + plan.rank * per) for r in range(plan.world_size)
                                    if r != plan.rank] for c in range(2)]
                    sh["barrier"] = plan.peer_barrier_args(sym[1])     # kernel-issued barrier: flag addresses, peers' ranks
                    sh["epoch"] = 0
                else:
                    sh["vecs"], sh["sym"], sh["peers"] = [torch.zeros(n, dtype=torch.float64, device=dev)] * 2, None, [[], []]
                sh["turn"] = 0
                sh["host"] = torch.zeros(per * plan.world_size, dtype=torch.float64).pin_memory()
            else:   # CPU collectives (gloo): the host-side logic of the same protocol
                sh["vecs"] = [torch.zeros(per * plan.world_size, dtype=torch.float64)] * 2
                sh["host"] = sh["vecs"][0]
                sh["turn"] = 0
            self._sh = sh
        return self._sh

    def _log_likelihood_sharded(self, dirty: List[int], gene_trees: Sequence[Any], flags: int) -> None:
        """This rank's dirty loci -> its slice of the device vector -> one in-place all-gather -> one D2H copy.
        Every rank runs the same chain, so every rank knows which loci are dirty anywhere (``_dirty_any``)."""
        import torch
        plan = self._shard
        lo, hi = plan.local_range()
        if not self._dirty_any and not dirty:
            return
        sh = self._shard_state()
        per = sh["per"]
        base = plan.rank * per
        turn = sh["turn"] = sh["turn"] ^ 1
        vec = sh["vecs"][turn]
        if sh["gpu"]:
            with torch.cuda.stream(sh["stream"]):
                fused = False
                if dirty:
                    if self.resident_batches and len(dirty) == hi - lo and _device_kind(self._model) != "explicit" \
                            and not (flags & _lib.EVAL_FULL):
                        # every rank takes this branch together (every rank's shard is dirty as a whole): with symmetric
                        # memory the kernel's peer stores + one device-side barrier replace the all-gather
                        fused = sh["sym"] is not None and len(self._dirty_any) == self.n_loci
                        if fused:
                            sh["epoch"] += 1      # every rank counts the same fused evaluations
                        self._eval_all_local(dirty, gene_trees, out_device_ptr=vec.data_ptr() + 8 * base,
                                             peer_ptrs=sh["peers"][turn] if fused else [],
                                             barrier=(sh["barrier"] + (sh["epoch"],)) if fused else None)
                    elif len(dirty) == hi - lo:
                        # the whole shard through pnb_eval (dirty-path aware, pending): the kernel writes this rank's slice
                        self.felsenstein_batch(dirty, gene_trees, flags=flags, out_device_ptr=vec.data_ptr() + 8 * base)
                        self._pending = list(dirty)
                    else:
                        vals = self.felsenstein_batch(dirty, gene_trees, flags=flags)
                        self._pending = list(dirty)
                        pos = torch.from_numpy(base + np.asarray(dirty, dtype=np.int64) - lo).to(vec.device)
                        vec.index_copy_(0, pos, torch.from_numpy(np.ascontiguousarray(vals)).to(vec.device))
                if not fused:    # (fused: the kernel's last block has exchanged epochs with every peer before it completed)
                    plan.allgather_slices_(vec)
                sh["host"].copy_(vec, non_blocking=True)
            sh["stream"].synchronize()
            if fused and self._batch is not None and self._batch.barrier_timed_out():
                raise RuntimeError("sharded evaluation: a peer GPU did not reach the barrier of the fused collective within 4 s")
        else:
            if dirty:
                vals = self.felsenstein_batch(dirty, gene_trees, flags=flags)
                self._pending = list(dirty)
                vec.numpy()[base + np.asarray(dirty, dtype=np.int64) - lo] = vals
            plan.allgather_slices_(vec)
        # only the entries that were dirty somewhere are taken from the gathered vector (a slot of the device vector
        # may still hold the value of a proposal that was rejected since)
        if len(self._dirty_any) == self.n_loci:
            self._felsen[:] = sh["host"].numpy()[sh["index"]].tolist()     # one C-level conversion, locus order
        else:
            hv, ix = sh["host"].numpy(), sh["index"]
            for i in self._dirty_any:
                self._felsen[i] = float(hv[ix[i]])
        self._dirty_any.clear()

    def _total(self, gene_trees: Sequence[Any], species_net: Any, theta: float) -> float:
        """The sum over loci once every ``_felsen`` entry is filled; fills the ``None`` entries of ``_msnc``."""
        if self._msnc_fn is None and species_net is None:
            total = _list_total(self._felsen)
        elif self._msnc_fn is None:
            self._fill_msnc(gene_trees, species_net, theta)
            # blocked sums instead of the reference's Python loop over loci (src/_mcmc_seq.py:745-761): at 10^3 loci that
            # loop costs more than the device work of a single-locus move (_SummedList)
            total = _list_total(self._felsen) + _list_total(self._msnc)
        else:
            total = 0.0
            for i in range(self.n_loci):  # reference loop, src/_mcmc_seq.py:745-761
                if self._msnc[i] is None:
                    self._msnc[i] = float(self._msnc_fn(i, gene_trees[i], species_net, theta))
                total += self._felsen[i] + self._msnc[i]
        if not math.isfinite(total):
            return float("-inf")
        return total

    # -- several chains at once (MC3) ----------------------------------------
    @staticmethod
    def log_likelihood_many(engines: Sequence["SeqLikelihoodEngine"], gene_trees: Sequence[Sequence[Any]],
                            species_nets: Optional[Sequence[Any]] = None, thetas: Optional[Sequence[float]] = None
                            ) -> List[float]:
        """``log_likelihood`` of several engines -- the chains of a Metropolis-coupled run
        (``src/_mcmc_seq.py:3584-3723``), or independent chains sharing one device -- with the dirty loci of ALL
        of them in ONE ``pnb_eval``: per chain a gene-tree move dirties one locus, so T chains cost one launch and
        one synchronisation instead of T.  Each engine keeps its own partials cache, pending / commit / rollback
        state and scalar caches, exactly as if it had been evaluated alone (results are bit-identical).  The
        engines must share context, model and numerical mode and be unsharded."""
        T = len(engines)
        if len(gene_trees) != T:
            raise ValueError("log_likelihood_many: one gene-tree list per engine")
        species_nets = [None] * T if species_nets is None else list(species_nets)
        thetas = [0.0] * T if thetas is None else list(thetas)
        if T == 0:
            return []
        e0 = engines[0]
        ctx = e0._context()
        for e in engines:
            if e._shard.world_size != 1:
                raise ValueError("log_likelihood_many: sharded engines are evaluated one by one")
            if e._context() is not ctx or e._model is not e0._model or e._mode != e0._mode:
                raise ValueError("log_likelihood_many: engines must share context, model and numerical mode")
        owners: List[Tuple["SeqLikelihoodEngine", List[int]]] = []
        flats: List[FlatTree] = []
        rows: List[dict] = []
        handles: List[np.ndarray] = []
        full = False
        for e, trees in zip(engines, gene_trees):
            e._commit_pending()
            dirty = sorted(i for i in e._dirty if e._felsen[i] is None)
            e._dirty.clear()
            full = full or e._full_next
            e._full_next = False
            owners.append((e, dirty))
            for i in dirty:
                flats.append(flatten(trees[i]))
                rows.append(e._calcs[i]._row_of)
            if dirty:
                handles.append(e._handles(dirty))
        if flats:
            node_off, parent, blen, tip_row = pack_trees(flats, rows)
            P_edge = _explicit_P_edges(e0._model, blen, 1.0) if _device_kind(e0._model) == "explicit" else None
            flags = _lib.EVAL_PENDING | e0._mode | (_lib.EVAL_FULL if full else 0)
            vals = ctx.eval(_model_handle(e0._model, ctx), np.concatenate(handles), node_off, parent, blen, tip_row,
                            1.0, flags, P_edge)
            stats = ctx.last_stats()
            k = 0
            for e, dirty in owners:
                for i in dirty:
                    e._felsen[i] = float(vals[k])
                    k += 1
                e._pending = list(dirty)
                e.last_stats = stats
        # every _felsen entry is filled and PENDING on the device; each engine adds its own MSNC factor and sums
        # (not through log_likelihood(): that would commit what has just been evaluated)
        out = []
        for e, trees, net, th in zip(engines, gene_trees, species_nets, thetas):
            try:
                out.append(e._total(trees, net, th))
            except TypeError as exc:
                raise RuntimeError("log_likelihood_many: a cache entry was set to None without invalidate_locus(); "
                                   "evaluate that engine with log_likelihood() once") from exc
        return out

    # -- MSNC density on the device ----------------------------------------------------------------
    def _network(self, species_net: Any) -> SpeciesNetwork:
        if species_net is not self._net_obj or self._net is None:
            new = as_species_network(species_net)
            # species ids are the ranks of the leaf labels: a height, gamma or topology move that keeps the species
            # (every move of the sampler does) keeps the per-locus id arrays and the packed gene-tree store
            if self._net is None or new.species != self._net.species:
                self._sp_ids = [None] * self.n_loci
                self._store = None
            elif self._store is not None:
                self._store["net"] = new
            self._net = new
            self._net_obj = species_net
        return self._net

    def _species_ids(self, i: int, ft: FlatTree, net: SpeciesNetwork) -> np.ndarray:
        hit = self._sp_ids[i]
        if hit is not None and hit[0] is ft.label:
            return hit[1]
        species_of = self._species_of
        if species_of is None:   # one allele per species, named after it
            species_of = {lab: lab for lab in ft.label if lab is not None}
        ids = net.species_ids_of(ft, species_of)
        self._sp_ids[i] = (ft.label, ids)
        return ids

    def _recount_msnc(self) -> None:
        none = [i for i, v in enumerate(self._msnc) if v is None]
        self._msnc_all = len(none) == self.n_loci
        self._msnc_dirty = set() if self._msnc_all else set(none)

    def _fill_msnc(self, gene_trees: Sequence[Any], species_net: Any, theta: float) -> None:
        """Every ``None`` entry of ``_msnc`` in one ``pnb_msnc_eval`` (all ranks evaluate all loci: the
        work is tiny next to the pruning and it saves a second collective).  Which entries are ``None`` is
        tracked, not searched; after ``invalidate_network`` the trees of ALL loci go to the device from a
        packed store that is updated in place for the loci whose tree object changed."""
        if not self._msnc_all and not self._msnc_dirty:
            return
        theta = _check_theta(theta)
        net = self._network(species_net)
        if self._msnc_all:
            need: Sequence[int] = range(self.n_loci)
            off, parent, blen, ids = self._packed_all(gene_trees, net)
        else:
            need = sorted(i for i in self._msnc_dirty if self._msnc[i] is None)
            if not need:
                self._msnc_dirty.clear()
                return
            flats = [flatten(gene_trees[i]) for i in need]
            if len(flats) == 1:
                t = flats[0]
                off = np.array([0, t.n_nodes], dtype=np.int64)
                parent, blen, ids = t.parent, t.blen, self._species_ids(need[0], t, net)
            else:
                sizes = np.fromiter((t.n_nodes for t in flats), dtype=np.int64, count=len(flats))
                off = np.zeros(len(flats) + 1, dtype=np.int64)
                np.cumsum(sizes, out=off[1:])
                parent = np.ascontiguousarray(np.concatenate([t.parent for t in flats]), dtype=np.int32)
                blen = np.ascontiguousarray(np.concatenate([t.blen for t in flats]), dtype=np.float64)
                ids = np.ascontiguousarray(np.concatenate([self._species_ids(i, t, net) for i, t in zip(need, flats)]),
                                           dtype=np.int32)
        vals = msnc_eval_packed(net, off, parent, blen, ids, theta, self._context(), self._branch_thetas)
        if self._msnc_all:
            self._msnc[:] = vals.tolist()   # in place: a caller may share this list
        else:
            for i, v in zip(need, vals.tolist()):
                self._msnc[i] = v
        self._msnc_all = False
        self._msnc_dirty.clear()

    def _packed_all(self, gene_trees: Sequence[Any], net: SpeciesNetwork):
        """CSR arrays of the current gene trees of all loci, kept between calls: only the rows of loci whose
        tree OBJECT changed are rewritten (the reference's identity convention, src/_mcmc_seq.py:682-694)."""
        st = self._store
        if st is not None and st["net"] is net and len(st["objs"]) == self.n_loci:
            objs, off = st["objs"], st["off"]
            ok = True
            for i, g in enumerate(gene_trees):
                if g is objs[i]:
                    continue
                t = flatten(g)
                a, b = int(off[i]), int(off[i + 1])
                if t.n_nodes != b - a:
                    ok = False
                    break
                st["parent"][a:b] = t.parent
                st["blen"][a:b] = t.blen
                st["ids"][a:b] = self._species_ids(i, t, net)
                objs[i] = g
            if ok:
                return off, st["parent"], st["blen"], st["ids"]
        flats = [flatten(g) for g in gene_trees]
        sizes = np.fromiter((t.n_nodes for t in flats), dtype=np.int64, count=len(flats))
        off = np.zeros(len(flats) + 1, dtype=np.int64)
        np.cumsum(sizes, out=off[1:])
        self._store = {
            "net": net, "objs": list(gene_trees), "off": off,
            "parent": np.ascontiguousarray(np.concatenate([t.parent for t in flats]), dtype=np.int32),
            "blen": np.ascontiguousarray(np.concatenate([t.blen for t in flats]), dtype=np.float64),
            "ids": np.ascontiguousarray(np.concatenate([self._species_ids(i, t, net) for i, t in enumerate(flats)]),
                                        dtype=np.int32),
        }
        st = self._store
        return off, st["parent"], st["blen"], st["ids"]

    def msnc_candidate_sets(self, sets: Sequence[Tuple[int, Any]], species_net: Any, theta: float) -> List[np.ndarray]:
        """log P(g | Psi) of the candidate trees of MANY loci in one device call -- the MSNC half of
        ``_enumerate_scored_candidates`` (reference ``src/_mcmc_seq.py:1935-1957``).  ``sets`` as for
        :meth:`score_candidate_sets`: ``[(locus, [tree, ...])]`` or ``[(locus, (base, parent[K, n], blen[K, n]))]``."""
        theta = _check_theta(theta)
        net = self._network(species_net)
        parents, blens, idss, counts, per_set = [], [], [], [], []
        for locus, trees in sets:
            if isinstance(trees, tuple) and len(trees) == 3 and isinstance(trees[1], np.ndarray):
                base, par, bl = trees
                K, n = par.shape
                parents.append(par.reshape(-1))
                blens.append(bl.reshape(-1))
                idss.append(np.tile(self._species_ids(locus, flatten(base), net), K))   # leaves stay leaves under NNI
                counts.append(np.full(K, n, dtype=np.int64))
                per_set.append(K)
            else:
                for t in trees:
                    ft = flatten(t)
                    parents.append(ft.parent)
                    blens.append(ft.blen)
                    idss.append(self._species_ids(locus, ft, net))
                    counts.append(np.array([ft.n_nodes], dtype=np.int64))
                per_set.append(len(trees))
        if not counts:
            return [np.zeros(0) for _ in sets]
        sizes = np.concatenate(counts)
        off = np.zeros(sizes.shape[0] + 1, dtype=np.int64)
        np.cumsum(sizes, out=off[1:])
        out = msnc_eval_packed(net, off, np.ascontiguousarray(np.concatenate(parents), dtype=np.int32),
                               np.ascontiguousarray(np.concatenate(blens), dtype=np.float64),
                               np.ascontiguousarray(np.concatenate(idss), dtype=np.int32), theta, self._context(),
                               self._branch_thetas)
        res, k = [], 0
        for K in per_set:
            res.append(out[k:k + K])
            k += K
        return res

    def network_only_msnc_score(self, gene_trees: Sequence[Any], species_net: Any, theta: float,
                                scales: Sequence[float]) -> float:
        """``sum_i logsumexp_s log P(g_i with heights scaled by s | candidate network)`` -- the likelihood part of the
        reference's ``_network_only_log_score`` (``src/_mcmc_seq.py:1163-1214``), the score the informed add / delete /
        relocate moves rank every candidate placement by.  The reference runs ``n_loci x len(scales)`` host densities
        per candidate (94 % of its chain time on 30 loci); here one candidate is ONE ``pnb_msnc_eval`` over the gene
        trees of all loci at all scales, packed once per set of gene-tree objects (they do not change inside a network
        move).  The candidate never touches the engine's own network / cache state."""
        theta = _check_theta(theta)
        cand = as_species_network(species_net)
        scales = tuple(float(x) for x in scales)
        pk = self._scan_pack
        if pk is None or pk["scales"] != scales or pk["species"] != cand.species or len(pk["objs"]) != len(gene_trees) \
                or any(a is not b for a, b in zip(pk["objs"], gene_trees)):
            species_of = self._species_of
            parents, blens, idss, sizes = [], [], [], []
            for g in gene_trees:
                ft = flatten(g)
                so = species_of if species_of is not None else {lab: lab for lab in ft.label if lab is not None}
                ids = cand.species_ids_of(ft, so)
                for sc in scales:
                    parents.append(ft.parent)
                    blens.append(ft.blen if sc == 1.0 else ft.blen * sc)   # powers of two: the scaled times are exact
                    idss.append(ids)
                    sizes.append(ft.n_nodes)
            off = np.zeros(len(sizes) + 1, dtype=np.int64)
            np.cumsum(np.asarray(sizes, dtype=np.int64), out=off[1:])
            pk = self._scan_pack = {
                "scales": scales, "species": cand.species, "objs": list(gene_trees), "off": off,
                "parent": np.ascontiguousarray(np.concatenate(parents), dtype=np.int32),
                "blen": np.ascontiguousarray(np.concatenate(blens), dtype=np.float64),
                "ids": np.ascontiguousarray(np.concatenate(idss), dtype=np.int32)}
        dens = msnc_eval_packed(cand, pk["off"], pk["parent"], pk["blen"], pk["ids"], theta, self._context(), self._branch_thetas)
        total = 0.0
        for row in dens.reshape(-1, len(scales)).tolist():     # _logsumexp over the finite terms (:1149-1160, :1209-1213)
            terms = [x for x in row if math.isfinite(x)]
            if not terms:
                return -math.inf
            m = max(terms)
            acc = 0.0
            for v in terms:
                acc += math.exp(v - m)
            total += m + math.log(acc)
        return total

    # -- natively enumerated candidate batches (candidates.CandidateBatch) ---------------------------------
    def score_batch(self, batch: Any, loci: Sequence[int], *, site_rate: float = 1.0) -> np.ndarray:
        """Felsenstein log-likelihood of every member of a :class:`~phynetpy_b200.candidates.CandidateBatch`
        (tree ``j`` of the batch belongs to locus ``loci[j]``) in ONE ``PNB_EVAL_NOSTORE`` call; the arrays the
        native enumerator wrote are handed to ``pnb_eval`` as they are."""
        self._commit_pending()
        ctx = self._context()
        counts = np.diff(batch.member_off)
        handles = np.repeat(self._handles(loci), counts)
        tip_row = batch.expand([batch.trees[j].tip_rows(self._calcs[i]._row_of) for j, i in enumerate(loci)], np.int32)
        P_edge = _explicit_P_edges(self._model, batch.blen, site_rate) if _device_kind(self._model) == "explicit" else None
        out = ctx.eval(_model_handle(self._model, ctx), handles, batch.node_off, batch.parent, batch.blen, tip_row,
                       site_rate, _lib.EVAL_NOSTORE | self._mode, P_edge)
        self._stats, self._stats_ctx = None, ctx      # fetched when somebody reads last_stats
        return out

    def msnc_batch(self, batch: Any, loci: Sequence[int], species_net: Any, theta: float) -> np.ndarray:
        """log P(g | Psi) of every member of a candidate batch in one ``pnb_msnc_eval`` call."""
        theta = _check_theta(theta)
        net = self._network(species_net)
        ids = batch.expand([self._species_ids(i, batch.trees[j], net) for j, i in enumerate(loci)], np.int32)
        return msnc_eval_packed(net, batch.node_off, batch.parent, batch.blen, ids, theta, self._context(), self._branch_thetas)

    def score_candidate_sets(self, sets: Sequence[Tuple[int, Sequence[Any]]], *, site_rate: float = 1.0) -> List[np.ndarray]:
        """Candidate trees for MANY loci in ONE launch: ``sets = [(locus, [tree, ...]), ...]``.

        This is the whole inner double loop of ``_coupled_gene_tree_reproposal`` (reference
        ``src/_mcmc_seq.py:2028`` over loci x ``_enumerate_scored_candidates`` ``:1935-1957`` over
        ``({g} U NNI(g)) x (0.25, 0.5, 1, 2, 4)``), i.e. up to 2*5*(1+2(n-2)) prunings per locus, as a
        single ``PNB_EVAL_NOSTORE`` batch.  Returns one array of log-likelihoods per set."""
        self._commit_pending()
        ctx = self._context()
        if sets and all(isinstance(t, tuple) and len(t) == 3 and isinstance(t[1], np.ndarray) for _, t in sets):
            return self._score_candidate_arrays(sets, site_rate)
        idx: List[int] = []
        flats: List[FlatTree] = []
        for locus, trees in sets:
            for t in trees:
                idx.append(locus)
                flats.append(flatten(t))
        if not idx:
            return [np.zeros(0) for _ in sets]
        node_off, parent, blen, tip_row = pack_trees(flats, [self._calcs[i]._row_of for i in idx])
        P_edge = _explicit_P_edges(self._model, blen, site_rate) if _device_kind(self._model) == "explicit" else None
        out = ctx.eval(_model_handle(self._model, ctx), self._handles(idx), node_off, parent, blen, tip_row,
                       site_rate, _lib.EVAL_NOSTORE | self._mode, P_edge)
        self._stats, self._stats_ctx = None, ctx      # fetched when somebody reads last_stats
        res, k = [], 0
        for _, trees in sets:
            res.append(out[k:k + len(trees)])
            k += len(trees)
        return res

    def _score_candidate_arrays(self, sets, site_rate: float) -> List[np.ndarray]:
        """``sets = [(locus, (base_tree, parent[K, n], blen[K, n])), ...]`` as produced by
        ``candidates.candidate_arrays``: packed with a handful of NumPy calls per locus instead of K tree
        objects (the host side of a 20,000-tree batch is otherwise slower than the GPU)."""
        ctx = self._context()
        handles, parents, blens, tips, counts = [], [], [], [], []
        for locus, (base, par, bl) in sets:
            K, n = par.shape
            calc = self._calcs[locus]
            handles.append(np.full(K, calc._locus_handle().value, dtype=np.uint64))
            parents.append(par.reshape(-1))
            blens.append(bl.reshape(-1))
            tips.append(np.tile(flatten(base).tip_rows(calc._row_of), K))   # leaves stay leaves under NNI
            counts.append(np.full(K, n, dtype=np.int64))
        sizes = np.concatenate(counts)
        node_off = np.zeros(sizes.shape[0] + 1, dtype=np.int64)
        np.cumsum(sizes, out=node_off[1:])
        parent = np.ascontiguousarray(np.concatenate(parents), dtype=np.int32)
        blen = np.ascontiguousarray(np.concatenate(blens), dtype=np.float64)
        tip_row = np.ascontiguousarray(np.concatenate(tips), dtype=np.int32)
        P_edge = _explicit_P_edges(self._model, blen, site_rate) if _device_kind(self._model) == "explicit" else None
        out = ctx.eval(_model_handle(self._model, ctx), np.concatenate(handles), node_off, parent, blen, tip_row,
                       site_rate, _lib.EVAL_NOSTORE | self._mode, P_edge)
        self._stats, self._stats_ctx = None, ctx      # fetched when somebody reads last_stats
        res, k = [], 0
        for _, (_, par, _) in sets:
            res.append(out[k:k + par.shape[0]])
            k += par.shape[0]
        return res

    def score_candidates(self, locus: int, trees: Sequence[Any], *, site_rate: float = 1.0) -> np.ndarray:
        """Felsenstein logL of many candidate trees for ONE locus in one launch (the
        loops of the coupled moves, reference ``src/_mcmc_seq.py:1864-1992``): cached
        partials of the committed tree are reused, nothing is written back."""
        self._commit_pending()
        idx = [locus] * len(trees)
        ctx = self._context()
        flats = [flatten(t) for t in trees]
        row = self._calcs[locus]._row_of
        node_off, parent, blen, tip_row = pack_trees(flats, [row] * len(trees))
        P_edge = _explicit_P_edges(self._model, blen, site_rate) if _device_kind(self._model) == "explicit" else None
        out = ctx.eval(_model_handle(self._model, ctx), self._handles(idx), node_off, parent, blen, tip_row,
                       site_rate, _lib.EVAL_NOSTORE | self._mode, P_edge)
        self._stats, self._stats_ctx = None, ctx      # fetched when somebody reads last_stats
        return out
