"""world_size-2 gloo test of the multi-GPU plumbing on CPU: the cyclic row shards, the slot layout that makes every
rank's rows one contiguous block, the all-gather of sampled factor blocks and the all-reduce of Normal-Wishart
statistics. The per-row arithmetic is the oracle's here (no GPU); what is under test is that the sharded sweep, stitched
together by the collectives exactly as DistributedSweep does on GPUs, equals the unsharded sweep bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bdf_b200.shard import MapShardPlan, ShardPlan, balanced_partition
from oracle import oracle as orc


def _problem():
    rng = np.random.default_rng(0)
    dims, D, nnz = [41, 23], 6, 700
    ids = np.stack([rng.integers(1, d + 1, nnz) for d in dims], axis=1)
    vals = rng.standard_normal(nnz)
    U = [rng.standard_normal((d, D)) * 0.3 for d in dims]
    Lambda = 2.0 * np.eye(D)
    mu = rng.standard_normal(D) * 0.1
    Z = [rng.standard_normal((d, D)) for d in dims]
    return dims, D, ids, vals, U, mu, Lambda, Z


def _plans(dims, ids, world, partition):
    if partition == "cyclic":
        return [ShardPlan(d, world) for d in dims]
    return [MapShardPlan(balanced_partition(np.bincount(ids[:, m] - 1, minlength=d), world, 5.0), world) for m, d in enumerate(dims)]


def _worker(rank, world, port, out, partition):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dims, D, ids, vals, U, mu, Lambda, Z = _problem()
    plans = _plans(dims, ids, world, partition)
    # every rank holds slot-ordered replicas of all factor matrices (like the device buffers)
    S = [torch.from_numpy(p.to_slots(u)) for p, u in zip(plans, U)]
    stats_all = []
    for mode in range(2):
        plan = plans[mode]
        Unat = [plans[m].from_slots(S[m].numpy()) for m in range(2)]
        # this rank samples only its rows: keep the observations of the rows it owns
        mine = plan.owner(ids[:, mode] - 1) == rank
        idf = orc.FastIDF(ids[mine], vals[mine], dims)
        Uw = [u.copy() for u in Unat]
        orc.sample_latent_all(idf, mode, Uw, 1.5, 0.2, mu, Lambda, Z[mode])
        blk = torch.zeros(plan.nper, D, dtype=torch.float64)
        rows = plan.local_rows(rank)
        blk[: len(rows)] = torch.from_numpy(Uw[mode][rows])
        dist.all_gather_into_tensor(S[mode], blk)            # equal-sized contiguous blocks
        # Normal-Wishart statistics of the local rows, then all-reduce of [N, NU, NS]
        n, NU, NS = orc.nw_stats(Uw[mode][rows])
        st = torch.from_numpy(np.concatenate([[n], NU, np.asarray(NS).ravel(order="F")]))
        dist.all_reduce(st)
        stats_all.append(st.numpy().copy())
    if rank == 0:
        np.savez(out, U0=plans[0].from_slots(S[0].numpy()), U1=plans[1].from_slots(S[1].numpy()), st0=stats_all[0], st1=stats_all[1])
    dist.destroy_process_group()


@pytest.mark.parametrize("partition", ["cyclic", "balanced"])
def test_sharded_sweep_equals_unsharded(tmp_path, partition):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, port, out, partition), nprocs=2, join=True)
    got = np.load(out)
    dims, D, ids, vals, U, mu, Lambda, Z = _problem()
    idf = orc.FastIDF(ids, vals, dims)
    Uo = [u.copy() for u in U]
    for mode in range(2):
        orc.sample_latent_all(idf, mode, Uo, 1.5, 0.2, mu, Lambda, Z[mode])
        n, NU, NS = orc.nw_stats(Uo[mode])
        st = got[f"st{mode}"]
        assert st[0] == n
        assert np.allclose(st[1:1 + D], NU, rtol=1e-13, atol=1e-13)
        assert np.allclose(st[1 + D:].reshape(D, D, order="F"), NS, rtol=1e-13, atol=1e-13)
    # rows with observations on both ranks are impossible (a row belongs to one rank), so the draws are bit-identical
    assert np.array_equal(got["U0"], Uo[0])
    assert np.array_equal(got["U1"], Uo[1])
