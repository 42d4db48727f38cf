"""C5 of BASELINE.json: sharded BPMF, synthetic 10M users × 1M items, 1B ratings, D=64, one process per GPU (8 × B200).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_c5.py [--scale S] [--steps K]

No rank ever holds the whole table (24 GB): the generator is counter-based (chunk c of the table is a pure function of (seed, c)),
every rank walks all chunks and keeps only the observations it needs — those whose user row OR item row it owns (the engine builds
the CSR of a mode from the observations whose row in that mode is local and ignores the rest) — about 2/N of the table per rank.
A first pass over the chunks counts the row degrees for the work-balanced shard maps. NOT yet run at full scale;
`shard_table` / `degrees` are covered by tests/test_host.py at small scale."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

N_USERS, N_ITEMS, NNZ, D = 10_000_000, 1_000_000, 1_000_000_000, 64
SEED = 20161017 + 4
CHUNK = 10_000_000
ALPHA = 1.5


def chunk(c, n1, n2, nnz, seed=SEED, skew=2.5, chunk_size=CHUNK):
    """Observations [c·chunk_size, …) of the table: 1-based (user, item) ids with skewed marginals and a value. Pure function of
    (seed, c): every rank regenerates the same chunk."""
    lo = c * chunk_size
    n = max(0, min(chunk_size, nnz - lo))
    rng = np.random.default_rng([seed, c])
    i1 = np.minimum((n1 * rng.random(n) ** skew).astype(np.int64), n1 - 1)
    i2 = np.minimum((n2 * rng.random(n) ** skew).astype(np.int64), n2 - 1)
    v = np.sin(0.001 * i1) * np.cos(0.01 * i2) + rng.standard_normal(n) / np.sqrt(ALPHA)  # a smooth planted signal + N(0, 1/alpha)
    return i1 + 1, i2 + 1, v


def degrees(n1, n2, nnz, **kw):
    d1, d2 = np.zeros(n1, dtype=np.int64), np.zeros(n2, dtype=np.int64)
    for c in range((nnz + kw.get("chunk_size", CHUNK) - 1) // kw.get("chunk_size", CHUNK)):
        i1, i2, _ = chunk(c, n1, n2, nnz, **kw)
        d1 += np.bincount(i1 - 1, minlength=n1)
        d2 += np.bincount(i2 - 1, minlength=n2)
    return d1, d2


def shard_table(rank, owner1, owner2, n1, n2, nnz, **kw):
    """The part of the table rank `rank` must register: observations whose user or item row it owns, in table order."""
    ids, vals = [], []
    for c in range((nnz + kw.get("chunk_size", CHUNK) - 1) // kw.get("chunk_size", CHUNK)):
        i1, i2, v = chunk(c, n1, n2, nnz, **kw)
        keep = (owner1[i1 - 1] == rank) | (owner2[i2 - 1] == rank)
        ids.append(np.stack([i1[keep], i2[keep]], axis=1))
        vals.append(v[keep])
    return np.concatenate(ids), np.concatenate(vals)


def main():
    import torch
    import torch.distributed as dist

    import bdf_b200
    from bdf_b200.shard import DistributedSweep, balanced_partition

    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n1, n2, nnz = int(N_USERS * args.scale), int(N_ITEMS * args.scale), int(NNZ * args.scale)
    d1, d2 = degrees(n1, n2, nnz)
    o1, o2 = balanced_partition(d1, world, 2.0 * D), balanced_partition(d2, world, 2.0 * D)
    ids, vals = shard_table(rank, o1, o2, n1, n2, nnz)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng = bdf_b200.Engine(D, device=local, rank=rank, world=world)
    eng.set_stream(stream.cuda_stream)
    eng.set_seed(SEED)
    e1, e2 = eng.add_entity_partitioned(n1, o1), eng.add_entity_partitioned(n2, o2)
    rel = eng.add_relation([e1, e2], ids, vals)
    eng.set_relation_params(rel, ALPHA, 0.0)
    del ids, vals
    ds = DistributedSweep(eng, [e1, e2])
    ds.sweep(args.warmup)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ds.sweep(args.steps); b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item()) / args.steps
        flops = 2 * nnz * (D * (D + 1) + 2 * D) + (n1 + n2) * (D ** 3 / 3 + 2 * D * D)
        print(json.dumps({"config": f"C5 sharded BPMF {n1}x{n2}, {nnz} ratings, D={D}", "n_gpus": world, "sweeps_per_s": 1e3 / ms, "ms_per_sweep": ms,
                          "algorithmic_tflops_total": flops / ms / 1e9}))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
