"""`macau(..., devices=[0, 1, 2, 3])` — the Gibbs loop on several GPUs through the reference's own entry point.

The reference selects its workers with a keyword (`latent_pids = workers()`, src/macau.jl:12) and deals the rows of the sampled mode out
to them (src/sampling.jl:149-172). Here the keyword is `devices`: one process per GPU (the calling process drives devices[0] and spawns
one worker per further device; under `torchrun`, where a process group already exists, every process simply takes devices[rank]). Every
process runs the SAME host loop (`macau._macau_loop`) on its own handle (`bdf_create(..., rank, world)`): it draws the rows it owns, the
row kernel stores each drawn row straight into every peer's replica over NVLink, and the small reductions — the (1 + D + D²)
Normal-Wishart statistics, the training SSE behind alpha, the test-set sums — are NCCL all-reduces issued from `Comm`. All ranks then draw
identical hyper-parameters (the same Philox key, or the same seeded host generator), so no rank ever waits for another's host. Rank 0
prints, writes the dumps and returns the result.
"""
from __future__ import annotations

import os
import socket

import numpy as np

from .engine import Engine
from .shard import balanced_partition, device_view


class Comm:
    """The collectives of one rank of a multi-GPU run (torch.distributed, NCCL over NVLink)."""

    def __init__(self, rank: int, world: int, device: int, num_latent: int):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank, self.world, self.device, self.D = rank, world, device, num_latent
        self.dev = torch.device("cuda", device)
        # ONE stream for the engine's kernels and the collectives: a non-default torch stream made current (the NULL handle of the default
        # stream would select the engine's own stream)
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        self.views = {}
        self.fused = False

    def add_entity(self, eng: Engine, en, data) -> int:
        """Rows are dealt to the ranks by a work-balanced map (every rank computes the same one): with heavy-tailed degrees the cyclic deal
        of src/sampling.jl:154 hands the heaviest row of every group of `world` to rank 0."""
        deg = np.zeros(en.count)
        for r in en.relations:
            m = next(i for i, e2 in enumerate(r.entities) if e2 is en)
            deg += np.bincount(r.data.ids[:, m] - 1, minlength=en.count)
        return eng.add_entity_partitioned(en.count, balanced_partition(deg, self.world, 2.0 * self.D))

    def connect(self, eng: Engine, ents, beta_ents=()):
        """Bind the engine to the collectives' stream and map every peer's factor replicas (CUDA IPC) so that the row kernel's peer stores
        replace the all-gather of the drawn rows; without peer access the ranks fall back to an NCCL all-gather per half-sweep."""
        torch, dist = self.torch, self.dist
        eng.set_stream(self.stream.cuda_stream)
        for e in ents:
            ptr, nper, ld = eng.factors_dev(e)
            sptr, scount = eng.stats_dev(e)
            self.views[e] = (device_view(ptr, (self.world * nper * ld,), self.dev), nper * ld, device_view(sptr, (scount,), self.dev))
        ok = 1
        if self.world <= 8:
            mine = {e: eng.ipc_export(e) for e in ents}
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine)
            try:
                for r, handles in enumerate(everyone):
                    if r != self.rank:
                        for e in ents:
                            eng.ipc_import(e, r, handles[e])
            except Exception as exc:
                ok = 0
                print(f"[bdf_b200] rank {self.rank}: peer mapping failed ({exc}); using the NCCL all-gather", flush=True)
        else:
            ok = 0
        flag = torch.tensor([ok], device=self.dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self.fused = bool(flag.item())
        self.eng = eng
        # link-matrix solves are split by right-hand-side column over the ranks (solve_cg2, src/parallel_matrix.jl:488-507) once every
        # rank has its peers' beta buffers mapped; all ranks must agree, so the mappings are attempted only if the factor mappings worked
        self.beta_split = False
        if self.fused and beta_ents:
            mine = {e: eng.ipc_export_beta(e) for e in beta_ents}
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine)
            for r, handles in enumerate(everyone):
                if r != self.rank:
                    for e in beta_ents:
                        eng.ipc_import_beta(e, r, handles[e])
            self.beta_split = True

    def nw_stats(self, eng: Engine, e: int, uhat: bool) -> None:
        """ConditionalNormalWishart's reductions over ALL rows: this rank's partial statistics on the device, then the all-reduce (which
        also orders every rank's peer stores before the next half-sweep reads them)."""
        U, blk, stats = self.views[e]
        if not self.fused:
            self.dist.all_gather_into_tensor(U, U[self.rank * blk:(self.rank + 1) * blk])
        if uhat:
            eng.nw_stats_uhat(e)
        else:
            eng.step_nw_stats(e)
        self.dist.all_reduce(stats)

    def allreduce_scalars(self, xs):
        t = self.torch.tensor([float(x) for x in xs], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t)
        return t.tolist()

    def gather_strided(self, local, n: int):
        """Every rank holds elements rank::world of a length-n vector; returns the assembled vector (on every rank)."""
        torch = self.torch
        per = (n + self.world - 1) // self.world
        mine = torch.zeros(per, device=self.dev, dtype=torch.float64)
        mine[: len(local)] = torch.from_numpy(np.ascontiguousarray(local)).to(self.dev)
        out = torch.zeros(per * self.world, device=self.dev, dtype=torch.float64)
        self.dist.all_gather_into_tensor(out, mine)
        out = out.cpu().numpy().reshape(self.world, per)
        full = np.zeros(n)
        for r in range(self.world):
            cnt = len(range(r, n, self.world))
            full[r::self.world] = out[r, :cnt]
        return full

    def fence(self):
        """Stream-ordered rendezvous without a host sync: no rank's next kernel starts before every rank has finished what it enqueued so
        far. Needed at the end of an iteration: rank 0 may still be READING the factor replicas (pred_all, sample dumps, RMSE_train, the
        refresh before f(data)) while the other ranks would already store the next iteration's rows into them."""
        if not hasattr(self, "_fence"):
            self._fence = self.torch.zeros(1, device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(self._fence)

    def barrier(self):
        self.dist.barrier()
        self.torch.cuda.synchronize()


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_rank(rank: int, world: int, devices, data, kw, own_group: bool, port: int):
    import torch
    import torch.distributed as dist

    from .macau import _macau_loop

    device = int(devices[rank])
    torch.cuda.set_device(device)
    if own_group:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", device))
    comm = Comm(rank, world, device, kw["num_latent"])
    eng = Engine(kw["num_latent"], device=device, rank=rank, world=world)
    try:
        res = _macau_loop(data, eng, comm, **kw)
        comm.barrier()
    finally:
        eng.close()
        if own_group:
            dist.destroy_process_group()
    return res


def _worker(rank, world, devices, data, kw, port):
    _run_rank(rank, world, devices, data, kw, True, port)


def macau_multi(data, devices, kw):
    """Entry from `macau(devices=[...])`. Under torchrun (a process group exists) this process is one rank of it; otherwise it becomes
    rank 0 and spawns the other ranks."""
    import torch
    import torch.distributed as dist

    world = len(devices)
    if any(r.hasFeatures() for r in data.relations):
        raise ValueError("relation-level features run on one GPU (devices=[d]): bdf_set_relation_features")
    if dist.is_available() and dist.is_initialized():
        if dist.get_world_size() != world:
            raise ValueError(f"devices has {world} entries but the process group has {dist.get_world_size()} ranks")
        return _run_rank(dist.get_rank(), world, devices, data, kw, False, 0)
    if torch.cuda.device_count() < world:
        raise ValueError(f"devices={devices}: only {torch.cuda.device_count()} CUDA devices are visible")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    port = _free_port()
    kw_workers = dict(kw, f=None)  # the callback runs on rank 0 only (and need not be picklable); everything else is identical on all ranks
    procs = [ctx.Process(target=_worker, args=(r, world, devices, data, kw_workers, port), daemon=True) for r in range(1, world)]
    for p in procs:
        p.start()
    try:
        res = _run_rank(0, world, devices, data, kw, True, port)
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.terminate()
    bad = [p.exitcode for p in procs if p.exitcode != 0]
    if bad:
        raise RuntimeError(f"worker processes of macau(devices=...) failed with exit codes {bad}")
    return res
