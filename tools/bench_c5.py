"""C5 of BASELINE.json: sharded BPMF, synthetic 10M users × 1M items, 1B ratings, D=64, one process per GPU (8 × B200).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_c5.py [--scale S] [--steps K]

No rank ever holds the whole table (24 GB): the generator is counter-based (chunk c of the table is a pure function of (seed, c)),
every rank walks all chunks and keeps only the observations it needs — those whose user row OR item row it owns (the engine builds
the CSR of a mode from the observations whose row in that mode is local and ignores the rest) — about 2/N of the table per rank.
A first pass over the chunks counts the row degrees for the work-balanced shard maps. On a GPU the chunks are generated on the device
(`chunk_dev`: torch's counter-based Philox generator seeded by (seed, c) — identical on every rank, seconds instead of minutes for 1B
observations); the numpy twin `chunk` / `degrees` / `shard_table` is what tests/test_host.py covers on CPU.
`--check` (small scales): rank 0 also runs the same table on ONE GPU and the factors after the sweeps must agree (N GPUs == 1 GPU).
Prints one JSON line: sweeps/s (device-resident, CUDA events, max over ranks), end-to-end sweeps/s through the host-buffer C-ABI
sequence, row-kernel time per mode and its fraction of the FP64 DMMA peak, per-rank HBM in use, NVLink bytes stored per half-sweep."""
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


def chunk_dev(c, n1, n2, nnz, dev, seed=SEED, skew=2.5, chunk_size=CHUNK):
    """`chunk` on the device. Same construction, torch's Philox stream instead of numpy's PCG64."""
    import torch

    lo = c * chunk_size
    n = max(0, min(chunk_size, nnz - lo))
    g = torch.Generator(device=dev)
    g.manual_seed(seed * 1_000_003 + c)
    u = torch.rand(2, n, generator=g, device=dev, dtype=torch.float64)
    i1 = torch.clamp((n1 * u[0] ** skew).long(), max=n1 - 1)
    i2 = torch.clamp((n2 * u[1] ** skew).long(), max=n2 - 1)
    v = torch.sin(0.001 * i1) * torch.cos(0.01 * i2) + torch.randn(n, generator=g, device=dev, dtype=torch.float64) / np.sqrt(ALPHA)
    return i1 + 1, i2 + 1, v


def degrees_dev(n1, n2, nnz, dev):
    import torch

    d1, d2 = torch.zeros(n1, dtype=torch.int64, device=dev), torch.zeros(n2, dtype=torch.int64, device=dev)
    for c in range((nnz + CHUNK - 1) // CHUNK):
        i1, i2, _ = chunk_dev(c, n1, n2, nnz, dev)
        d1 += torch.bincount(i1 - 1, minlength=n1)
        d2 += torch.bincount(i2 - 1, minlength=n2)
    return d1.cpu().numpy(), d2.cpu().numpy()


def shard_table_dev(rank, owner1, owner2, n1, n2, nnz, dev):
    """`shard_table` on the device; rank < 0 keeps everything (the one-GPU twin of --check)."""
    import torch

    o1, o2 = torch.from_numpy(owner1).to(dev), torch.from_numpy(owner2).to(dev)
    ids1, ids2, vals = [], [], []
    for c in range((nnz + CHUNK - 1) // CHUNK):
        i1, i2, v = chunk_dev(c, n1, n2, nnz, dev)
        if rank >= 0:
            keep = (o1[i1 - 1] == rank) | (o2[i2 - 1] == rank)
            i1, i2, v = i1[keep], i2[keep], v[keep]
        ids1.append(i1.cpu()); ids2.append(i2.cpu()); vals.append(v.cpu())
    ids = np.empty((sum(len(x) for x in ids1), 2), dtype=np.int64, order="F")
    ids[:, 0] = torch.cat(ids1).numpy()
    ids[:, 1] = torch.cat(ids2).numpy()
    return ids, torch.cat(vals).numpy()


def nvlink_bytes(index):
    """Cumulative NVLink data counters of one GPU (nvidia-smi nvlink -gt d): (tx_bytes, rx_bytes) summed over its links, or None."""
    import re
    import subprocess

    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True, timeout=30).stdout
        tx = sum(int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
        rx = sum(int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out))
        return (tx * 1024, rx * 1024) if (tx or rx) else None
    except Exception:
        return None


def main():
    import time

    import torch
    import torch.distributed as dist

    import bdf_b200
    from bdf_b200.shard import DistributedSweep, balanced_partition

    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--check", action="store_true", help="N GPUs == 1 GPU on the same table (small scales only)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n1, n2, nnz = int(N_USERS * args.scale), int(N_ITEMS * args.scale), int(NNZ * args.scale)
    t_setup = time.perf_counter()
    d1, d2 = degrees_dev(n1, n2, nnz, dev)
    o1, o2 = balanced_partition(d1, world, 2.0 * D), balanced_partition(d2, world, 2.0 * D)
    ids, vals = shard_table_dev(rank, o1, o2, n1, n2, nnz, dev)
    torch.cuda.empty_cache()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng = bdf_b200.Engine(D, device=local, rank=rank, world=world)
    eng.set_stream(stream.cuda_stream)
    eng.set_seed(SEED)
    e1, e2 = eng.add_entity_partitioned(n1, o1), eng.add_entity_partitioned(n2, o2)
    nloc = int(ids.shape[0])
    rel = eng.add_relation([e1, e2], ids, vals)
    eng.set_relation_params(rel, ALPHA, 0.0)
    ntest = min(1_000_000, nloc // 100)
    test_ids, test_vals = np.asfortranarray(ids[:ntest]), vals[:ntest].copy()   # (held-out metrics only exercise the path: these are training rows)
    del ids, vals
    ds = DistributedSweep(eng, [e1, e2])
    t_setup = time.perf_counter() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ds.sweep(args.warmup)
    barrier()
    nv0 = nvlink_bytes(local) if rank == 0 else None  # N/A on pools that do not expose the NVLink counters; then the computed figure stands alone
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ds.sweep(args.steps); b.record()
    barrier()
    nv1 = nvlink_bytes(local) if rank == 0 else None
    t = torch.tensor([a.elapsed_time(b)], device="cuda", dtype=torch.float64)
    if args.check:  # N GPUs == 1 GPU: the same number of device-resident sweeps on one handle that holds the whole table
        multi = [eng.get_factors(e1), eng.get_factors(e2)]
        eng.close()
        barrier()
        if rank == 0:
            ids, vals = shard_table_dev(-1, o1, o2, n1, n2, nnz, dev)
            one = bdf_b200.Engine(D, device=local)
            one.set_seed(SEED)
            f1, f2 = one.add_entity(n1), one.add_entity(n2)
            r1 = one.add_relation([f1, f2], ids, vals)
            one.set_relation_params(r1, ALPHA, 0.0)
            one.sweep(args.warmup + args.steps)
            errs = [float(np.max(np.abs(m - s)) / np.max(np.abs(s))) for m, s in zip(multi, [one.get_factors(f1), one.get_factors(f2)])]
            one.close()
            print(json.dumps({"config": f"C5 shape at scale {args.scale}: {n1}x{n2}, {nnz} ratings, D={D}", "n_gpus": world, "sweeps": args.warmup + args.steps,
                              "rel_err_vs_one_gpu": {"users": errs[0], "items": errs[1]}, "ms_per_sweep": float(t.item()) / args.steps}))
            assert max(errs) < 1e-7, errs
            print("C5 CHECK OK")
        barrier()
        if world > 1:
            dist.destroy_process_group()
        return
    # the row kernel alone, per mode
    kt = {e1: [], e2: []}
    for _ in range(3):
        for e in (e1, e2):
            ds.join()
            x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x.record(); eng.step_sample(e); y.record(); y.synchronize()
            kt[e].append(x.elapsed_time(y))
            if world > 1:
                ds.half_sweep(e)
    tk = torch.tensor([float(np.mean(kt[e1][1:])), float(np.mean(kt[e2][1:]))], device="cuda", dtype=torch.float64)
    free, total = torch.cuda.mem_get_info()
    mem = torch.tensor([float(total - free)], device="cuda", dtype=torch.float64)
    # end to end: the host-buffer call sequence of macau() (see bench.py)
    hyper = {e: (np.zeros(D), 5.0 * np.eye(D)) for e in (e1, e2)}
    eng.set_test(rel, test_ids, test_vals, None, 0.0)
    eng.set_async(True)

    def host_sweep():
        for e in (e1, e2):
            eng.sample_mode(e, hyper[e][0], hyper[e][1], None)
            eng.step_nw_stats(e)
            if world > 1:
                dist.all_reduce(ds.views[e][2])
            eng.nw_sample_async(e, np.zeros(D), 2.0, np.eye(D), float(D))
        for e in (e1, e2):
            hyper[e] = eng.nw_sample_fetch(e)
        eng.advance_sweep()
        return eng.predict_accumulate(rel, True)

    host_sweep()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_sweep()
    barrier()
    te = torch.tensor([(time.perf_counter() - t0) / args.steps], device="cuda", dtype=torch.float64)
    if world > 1:
        for x in (t, tk, mem, te):
            dist.all_reduce(x, op=dist.ReduceOp.MAX)
    eng.close()
    if rank == 0:
        ms = float(t.item()) / args.steps
        fl_mode = [nnz * (D * (D + 1) + 2 * D) + n * (D ** 3 / 3 + 2 * D * D) for n in (n1, n2)]
        flops = sum(fl_mode)
        try:
            peak = json.load(open("profiles/fp64_peak_r01.json"))["dmma_sustained_tflops"]
        except Exception:
            peak = 37.07
        ld = (D + 3) // 4 * 4
        line = {"config": f"C5 sharded BPMF {n1}x{n2}, {nnz} ratings, D={D}", "n_gpus": world, "sweeps_per_s": 1e3 / ms, "ms_per_sweep": ms,
                "e2e_sweeps_per_s": 1.0 / float(te.item()), "algorithmic_tflops_total": flops / ms / 1e9,
                "row_kernel_ms_per_mode_max_over_ranks": {"users": float(tk[0].item()), "items": float(tk[1].item())},
                "row_kernel_frac_of_fp64_dmma_peak": (flops / world) / ((float(tk[0].item()) + float(tk[1].item())) / 1e3) / 1e12 / peak,
                "whole_sweep_frac_of_fp64_dmma_peak": flops / world / (ms / 1e3) / 1e12 / peak,
                "hbm_in_use_gb_max_over_ranks": float(mem.item()) / 1e9, "observations_registered_per_rank": nloc,
                "nvlink_bytes_stored_per_rank_per_sweep": int((n1 + n2) / world * ld * 8 * (world - 1)),
                "nvlink_counters_gpu0_per_sweep": ({"tx_bytes": (nv1[0] - nv0[0]) / args.steps, "rx_bytes": (nv1[1] - nv0[1]) / args.steps,
                                                    "source": "nvidia-smi nvlink -gt d around the timed sweeps (row-kernel peer stores + the NCCL all-reduces of the statistics)"}
                                                   if nv0 and nv1 else None),
                "setup_seconds": t_setup, "partition": "work-balanced (LPT on the 200k heaviest rows, snake deal of the rest)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
