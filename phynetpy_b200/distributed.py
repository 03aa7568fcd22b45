"""Loci sharding across GPUs: one process per GPU, ``torch.distributed`` plumbing.

The reference has no multi-device path (its only parallelism is independent
chains in spawned processes, ``src/_mcmc_seq.py:4076-4274``).  Loci are
independent terms of a plain sum (``src/_mcmc_seq.py:746-761``), so they shard
without any data-path exchange; the single collective per evaluation is an
all-reduce (sum) of the M-long per-locus log-likelihood vector in which every
rank has filled only its own range (zeros elsewhere).  Adding zeros is exact, so
the reduced vector is bit-identical for any number of ranks, and the host sums it
in locus order.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np


def balanced_ranges(work: Sequence[float], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous locus ranges with near-equal total work (work_i ~ patterns_i * internal nodes_i)."""
    w = np.asarray(work, dtype=np.float64)
    M = w.shape[0]
    if world_size <= 1:
        return [(0, M)]
    cum = np.concatenate([[0.0], np.cumsum(w)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        k = int(np.searchsorted(cum, target, side="left"))
        # choose the boundary nearest to the target, keeping ranges monotone
        if k > 0 and abs(cum[k - 1] - target) <= abs(cum[min(k, M)] - target):
            k -= 1
        cuts.append(min(max(k, cuts[-1]), M))
    cuts.append(M)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


@dataclass
class ShardPlan:
    """Which loci this rank owns, and how the per-locus vector is combined."""

    n_loci: int
    rank: int
    world_size: int
    ranges: List[Tuple[int, int]]
    group: Optional[object] = None  # torch.distributed process group (None = default)

    @classmethod
    def single(cls, n_loci: int) -> "ShardPlan":
        return cls(n_loci, 0, 1, [(0, n_loci)])

    @classmethod
    def from_env(cls, n_loci: int, work: Optional[Sequence[float]] = None, group=None) -> "ShardPlan":
        """Rank / world size from an initialised ``torch.distributed`` (else single)."""
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(group), dist.get_world_size(group)
            else:
                rank, world = 0, 1
        except ImportError:
            rank, world = 0, 1
        w = np.ones(n_loci) if work is None else work
        return cls(n_loci, rank, world, balanced_ranges(w, world), group)

    def local_range(self) -> Tuple[int, int]:
        return self.ranges[self.rank]

    def owner_of(self, locus: int) -> int:
        for r, (lo, hi) in enumerate(self.ranges):
            if lo <= locus < hi:
                return r
        raise IndexError(locus)

    def allreduce_sum(self, vec: np.ndarray) -> np.ndarray:
        """Sum an M-long host vector over ranks (gloo on CPU tensors, NCCL via a staging
        CUDA tensor).  Exact when every entry is non-zero on at most one rank."""
        if self.world_size == 1:
            return vec
        import torch
        import torch.distributed as dist
        backend = dist.get_backend(self.group)
        t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64))
        if backend == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    # -- the per-locus vector as equal slices: rank r owns [r * per, (r + 1) * per) -------------------------
    @property
    def slice_len(self) -> int:
        """Length of one rank's slice of the gathered vector (the longest shard; shorter shards are padded)."""
        return max(hi - lo for lo, hi in self.ranges)

    def gathered_index(self) -> np.ndarray:
        """For every locus its position in the gathered vector (``world_size * slice_len`` long)."""
        per = self.slice_len
        idx = np.empty(self.n_loci, dtype=np.int64)
        for r, (lo, hi) in enumerate(self.ranges):
            idx[lo:hi] = r * per + np.arange(hi - lo)
        return idx

    # -- peer-mapped vector: the pruning kernel stores results into every peer's copy (fused collective) -------
    def symmetric_vector(self, copies: int = 2):
        """``copies`` x (``world_size * slice_len``) float64 CUDA values allocated in symmetric memory, with every peer's copy
        mapped into this process (``torch.distributed._symmetric_memory``: CUDA VMM handles exchanged through the
        process group's store).  Returns ``(tensor, handle)`` or ``None`` when the platform cannot do it (then the
        NCCL all-gather is used).  ``handle.buffer_ptrs[r]`` is rank r's vector, ``handle.barrier()`` a device-side
        barrier over NVLink signal pads on the current stream.  Two copies, used alternately by consecutive
        evaluations, keep a fast rank's next kernel from storing into a vector a slower rank is still reading
        (a rank cannot run two evaluations ahead: it waits in the barrier of the one in between)."""
        if self.world_size == 1 or self.world_size > 8:
            return None
        try:
            import torch
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            if dist.get_backend(self.group) != "nccl":
                return None
            # + BARRIER_SLOTS 8-byte words behind the vectors: the flag array of the kernel-issued barrier (peer_barrier_args)
            t = symm_mem.empty(copies * self.slice_len * self.world_size + self.BARRIER_SLOTS, dtype=torch.float64,
                               device=torch.device("cuda", torch.cuda.current_device()))
            t.zero_()
            hdl = symm_mem.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
            ok = torch.tensor([1.0 if len(hdl.buffer_ptrs) == self.world_size else 0.0], device=t.device)
        except Exception:
            try:
                import torch
                import torch.distributed as dist
                ok = torch.tensor([0.0], device=torch.device("cuda", torch.cuda.current_device()))
                t = hdl = None
            except Exception:
                return None
        import torch.distributed as dist
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)     # all ranks or none
        if ok.item() < 0.5:
            return None
        return t, hdl

    BARRIER_SLOTS = 16   # uint64 flags, one per rank (world_size <= 8), behind the vectors of symmetric_vector()

    def peer_barrier_args(self, handle, copies: int = 2):
        """What ``ResidentBatch.set_peer_barrier`` needs, from the handle ``symmetric_vector`` returned: the address (valid in
        this process) of slot [this rank] in every peer's flag array, the address of this rank's own array, and the peers'
        ranks.  The flags are zero at allocation; the caller counts epochs 1, 2, 3, ... -- one per sharded evaluation, the
        same on every rank."""
        off = 8 * copies * self.slice_len * self.world_size
        peers = [r for r in range(self.world_size) if r != self.rank]
        return ([int(handle.buffer_ptrs[r]) + off + 8 * self.rank for r in peers],
                int(handle.buffer_ptrs[self.rank]) + off, peers)

    def allgather_slices_(self, vec) -> None:
        """In-place all-gather of equal slices: every rank has written its own slice of ``vec``
        (``world_size * slice_len`` doubles; a CUDA tensor the pruning kernel has just filled, on the
        stream the kernel ran on); afterwards every rank holds all slices.  NCCL: one ``ncclAllGather``
        whose send buffer IS this rank's slice of the receive buffer -- no staging copy, no zero-fill, no
        host round trip.  Adding nothing to anything, the result is bit-identical for any number of ranks."""
        if self.world_size == 1:
            return
        import torch.distributed as dist
        per = self.slice_len
        mine = vec[self.rank * per:(self.rank + 1) * per]
        if dist.get_backend(self.group) == "nccl":
            dist.all_gather_into_tensor(vec, mine, group=self.group)
        else:   # gloo (CPU tests): gather into views of the same vector
            parts = list(vec.view(self.world_size, per).unbind(0))
            dist.all_gather(parts, mine.clone(), group=self.group)
