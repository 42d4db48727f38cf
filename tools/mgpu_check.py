"""torchrun --nproc-per-node N tools/mgpu_check.py — the N-GPU sharded sweep must reproduce the 1-GPU sweep (Philox noise is
keyed by global row id, collectives only move data): factors after 3 sweeps compared on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import bdf_b200
from bdf_b200.shard import DistributedSweep

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
D = int(sys.argv[1]) if len(sys.argv) > 1 else 32
rng = np.random.default_rng(0)
n1, n2, nnz = 5003, 1201, 200000
ids = np.stack([np.minimum((n1 * rng.random(nnz) ** 2.5).astype(np.int64), n1 - 1) + 1, rng.integers(1, n2 + 1, nnz)], axis=1)
vals = rng.standard_normal(nnz)


def run(r, w, fused=True, balanced=False):
    eng = bdf_b200.Engine(D, device=local, rank=r, world=w)
    eng.set_stream(stream.cuda_stream)
    eng.set_seed(7)
    if balanced and w > 1:
        from bdf_b200.shard import balanced_partition

        e1 = eng.add_entity_partitioned(n1, balanced_partition(np.bincount(ids[:, 0] - 1, minlength=n1), w, 2.0 * D))
        e2 = eng.add_entity_partitioned(n2, balanced_partition(np.bincount(ids[:, 1] - 1, minlength=n2), w, 2.0 * D))
    else:
        e1, e2 = eng.add_entity(n1), eng.add_entity(n2)
    rel = eng.add_relation([e1, e2], ids, vals)
    eng.set_relation_params(rel, 1.5, float(vals.mean()))
    ds = DistributedSweep(eng, [e1, e2], fused_allgather=fused) if w > 1 else None
    if w > 1:
        ds.sweep(3)
    else:
        eng.sweep(3)
    torch.cuda.synchronize()
    out = [eng.get_factors(e1), eng.get_factors(e2), eng.get_hyper(e1), eng.get_hyper(e2)]
    eng.close()
    return out


multi = run(rank, world, fused=True)
dist.barrier()
multi_nccl = run(rank, world, fused=False)
dist.barrier()
multi_bal = run(rank, world, fused=True, balanced=True)
dist.barrier()
if rank == 0:
    single = run(0, 1)
    errs = [float(np.max(np.abs(a - b)) / np.max(np.abs(b))) for a, b in zip(multi[:2], single[:2])]
    errs += [float(np.max(np.abs(multi[k][1] - single[k][1])) / np.max(np.abs(single[k][1]))) for k in (2, 3)]
    print(f"world={world} D={D}: rel err of factors/hyper vs 1-GPU run:", errs)
    assert max(errs) < 1e-9, errs
    errs_b = [float(np.max(np.abs(multi_bal[k] - single[k])) / np.max(np.abs(single[k]))) for k in (0, 1)]
    print(f"   work-balanced shard map vs 1-GPU run:", errs_b)
    assert max(errs_b) < 1e-9, errs_b
    assert all(np.array_equal(a, b) for a, b in zip(multi[:2], multi_nccl[:2])), "fused all-gather differs from the NCCL all-gather"
    print("MGPU OK (fused peer-store all-gather == NCCL all-gather == 1 GPU)")
dist.barrier()
dist.destroy_process_group()
