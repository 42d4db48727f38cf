"""Multi-GPU plumbing: one process per GPU, rows of every entity dealt cyclically to ranks (row i → rank i % world, the
reference's worker shards `i:Nprocs:N`, src/sampling.jl:154). Each half-sweep a rank samples only its own rows; the
collectives are torch.distributed's (NCCL over NVLink on GPUs, gloo in the CPU tests):

  all-gather of the freshly sampled factor slice  — replaces the per-half-sweep broadcast of `sample_m`, src/sampling.jl:165
  all-reduce of the (1 + D + D²) Normal-Wishart statistics — the reductions of src/sampling.jl:117-119

The device buffers belong to libbdf_b200.so; they are wrapped zero-copy as torch tensors through the CUDA array interface.
`ShardPlan` is pure host logic (which rank owns which row, where a row lives in the slot-ordered buffer) and is what the
gloo tests exercise together with the collective layout.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class ShardPlan:
    """Slot layout of an entity with `count` rows over `world` ranks: row i (0-based) lives at slot
    (i % world) * nper + i // world, so every rank's rows are one contiguous block of `nper` slots (the last block may
    be partly padding) and an all-gather of equal-sized blocks assembles the whole factor matrix."""
    count: int
    world: int

    @property
    def nper(self) -> int:
        return (self.count + self.world - 1) // self.world

    def owner(self, i):
        return np.asarray(i) % self.world

    def slot(self, i):
        i = np.asarray(i)
        return (i % self.world) * self.nper + i // self.world

    def nlocal(self, rank: int) -> int:
        return max(0, (self.count - rank + self.world - 1) // self.world)

    def local_rows(self, rank: int):
        return np.arange(rank, self.count, self.world)

    def to_slots(self, U):
        """(count, D) row-ordered matrix → (world*nper, D) slot-ordered buffer (padding rows zero)."""
        U = np.asarray(U)
        out = np.zeros((self.world * self.nper,) + U.shape[1:], dtype=U.dtype)
        out[self.slot(np.arange(self.count))] = U
        return out

    def from_slots(self, S):
        return np.asarray(S)[self.slot(np.arange(self.count))]


class MapShardPlan:
    """Slot layout of an entity created with `bdf_add_entity_partitioned`: an explicit `rank_of_row` map (e.g. from
    `balanced_partition`). Rows keep their relative order inside a shard, every shard is padded to the largest one, rank r's rows
    occupy slots [r·nper, r·nper + nlocal(r)) — the host mirror of the device's slot_of_row table, same interface as ShardPlan."""

    def __init__(self, rank_of_row, world: int):
        self.rank_of_row = np.asarray(rank_of_row, dtype=np.int64)
        self.count, self.world = int(self.rank_of_row.shape[0]), int(world)
        if self.count and (self.rank_of_row.min() < 0 or self.rank_of_row.max() >= world):
            raise ValueError("rank_of_row entries must lie in 0..world-1")
        cnt = np.bincount(self.rank_of_row, minlength=world)
        self.nper = int(max(1, cnt.max())) if self.count else 1
        self._cnt = cnt
        pos = np.zeros(self.count, dtype=np.int64)
        for r in range(world):
            idx = np.flatnonzero(self.rank_of_row == r)
            pos[idx] = np.arange(idx.shape[0])
        self._slot = self.rank_of_row * self.nper + pos

    def owner(self, i):
        return self.rank_of_row[np.asarray(i)]

    def slot(self, i):
        return self._slot[np.asarray(i)]

    def nlocal(self, rank: int) -> int:
        return int(self._cnt[rank])

    def local_rows(self, rank: int):
        return np.flatnonzero(self.rank_of_row == rank)

    def to_slots(self, U):
        U = np.asarray(U)
        out = np.zeros((self.world * self.nper,) + U.shape[1:], dtype=U.dtype)
        out[self._slot] = U
        return out

    def from_slots(self, S):
        return np.asarray(S)[self._slot]


class _DevArray:
    """Minimal __cuda_array_interface__ holder so torch can view a raw device pointer without copying."""

    def __init__(self, ptr: int, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}


def device_view(ptr: int, shape, device):
    import torch

    return torch.as_tensor(_DevArray(ptr, shape), device=device)


def balanced_partition(degrees, world: int, row_cost: float = 0.0, exact_top: int = 200_000):
    """Work-balanced shard map: rank_of_row for `world` ranks by greedy longest-processing-time assignment of the rows, heaviest
    first, each to the currently lightest rank. The weight of a row is its number of observations plus `row_cost` (the fixed
    factorisation + draw cost of a row in observation-equivalents). With a heavy-tailed degree distribution the cyclic deal
    `i:Nprocs:N` (src/sampling.jl:154) always hands the heaviest row of every group of `world` to the same rank.
    Entities with more than `exact_top` rows (C5: 10M users) get the exact greedy assignment for their `exact_top` heaviest rows and a
    vectorised snake deal (0 … W-1, W-1 … 0) of the remaining, light rows in descending weight order — each snake round trip adds
    the same load to every rank up to the weight difference inside it."""
    import heapq

    deg = np.asarray(degrees, dtype=np.float64) + float(row_cost)
    n = deg.shape[0]
    rank = np.zeros(n, dtype=np.int32)
    if world == 1:
        return rank
    order = np.argsort(-deg, kind="stable")
    head = order[:exact_top] if n > exact_top else order
    heap = [(0.0, r) for r in range(world)]
    for i in head.tolist():
        load, r = heap[0]
        rank[i] = r
        heapq.heapreplace(heap, (load + deg[i], r))
    if n > exact_top:
        tail = order[exact_top:]
        by_load = np.asarray([r for _, r in sorted(heap)], dtype=np.int32)   # lightest rank first
        pos = np.arange(tail.shape[0]) % (2 * world)
        snake = np.where(pos < world, pos, 2 * world - 1 - pos)
        rank[tail] = by_load[snake]
    return rank


class DistributedSweep:
    """Drives device-resident Gibbs sweeps over `world` GPUs: the loop body of src/macau.jl:96-134 with the collectives
    between the kernels. With world == 1 it degenerates to the same kernel sequence without communication."""

    def __init__(self, engine, entities, group=None, fused_allgather=True, overlap_draw=True):
        import torch
        import torch.distributed as dist

        self.eng = engine
        self.entities = list(entities)
        self.world = engine.world
        self.group = group
        self.dist = dist if engine.world > 1 else None
        dev = torch.device("cuda", torch.cuda.current_device())
        # ONE stream for the engine's kernels, the collectives and the side-stream events: bind the engine to torch's current stream.
        # The default stream's handle is NULL, which bdf_set_stream reads as "the engine's own stream", so in that case a fresh torch
        # stream is made current first. half_sweep() refuses to run if the caller has switched streams since.
        if torch.cuda.current_stream().cuda_stream == 0:
            torch.cuda.set_stream(torch.cuda.Stream())
        self.stream = torch.cuda.current_stream()
        engine.set_stream(self.stream.cuda_stream)
        self.views = {}
        for e in self.entities:
            ptr, nper, ld = engine.factors_dev(e)
            sptr, scount = engine.stats_dev(e)
            self.views[e] = (device_view(ptr, (engine.world * nper * ld,), dev), nper * ld, device_view(sptr, (scount,), dev))
        # fused all-gather: exchange CUDA IPC handles once; afterwards the row kernel writes every drawn row into all peer
        # replicas over NVLink and the per-half-sweep all-gather disappears
        self.fused = bool(fused_allgather) and engine.world > 1 and engine.world <= 8
        if self.fused:
            mine = {e: engine.ipc_export(e) for e in self.entities}
            everyone = [None] * engine.world
            dist.all_gather_object(everyone, mine, group=group)
            ok = 1
            try:
                for r, handles in enumerate(everyone):
                    if r != engine.rank:
                        for e in self.entities:
                            engine.ipc_import(e, r, handles[e])
            except Exception as exc:  # no peer access between some pair of GPUs: every rank falls back to the NCCL all-gather
                ok = 0
                print(f"[bdf_b200] rank {engine.rank}: peer mapping failed ({exc}); using the NCCL all-gather", flush=True)
            flag = torch.tensor([ok], device=dev, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            # (replicas already mapped on some ranks keep receiving peer stores; the all-gather then rewrites the same values)
            self.fused = bool(flag.item())

        # The Normal-Wishart draw of an entity is first needed by that entity's NEXT half-sweep, so it runs on a high-priority side
        # stream beside the following entity's row kernel (same kernels, same Philox streams: results are unchanged)
        self.overlap = bool(overlap_draw)
        self.side = torch.cuda.Stream(priority=-1) if self.overlap else None
        self.draw_done = {}

    def half_sweep(self, e):
        import torch

        eng = self.eng
        main = torch.cuda.current_stream()
        if main.cuda_stream != self.stream.cuda_stream:
            raise RuntimeError("DistributedSweep: the current CUDA stream changed since construction; the engine's kernels and the collectives must share one stream")
        if e in self.draw_done:
            main.wait_event(self.draw_done.pop(e))  # this entity's (mu, Lambda) from its previous draw
        eng.step_sample(e)
        U, blk, stats = self.views[e]
        if self.dist is not None and not self.fused:
            mine = U[eng.rank * blk:(eng.rank + 1) * blk]
            self.dist.all_gather_into_tensor(U, mine, group=self.group)
        eng.step_nw_stats(e)
        if self.dist is not None:
            self.dist.all_reduce(stats, group=self.group)
        if not self.overlap:
            eng.step_nw_draw(e)
            return
        ready = torch.cuda.Event()
        ready.record(main)
        self.side.wait_event(ready)
        eng.step_nw_draw_on(e, self.side.cuda_stream)
        done = torch.cuda.Event()
        done.record(self.side)
        self.draw_done[e] = done

    def join(self):
        """Order every outstanding draw before whatever the caller enqueues next on the current stream."""
        import torch

        if not self.draw_done:
            return
        main = torch.cuda.current_stream()
        for e in list(self.draw_done):
            main.wait_event(self.draw_done.pop(e))

    def sweep(self, n: int = 1):
        for _ in range(n):
            for e in self.entities:
                self.half_sweep(e)
            self.eng.advance_sweep()
        if self.overlap:
            self.join()
