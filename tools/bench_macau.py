"""C3 of BASELINE.json: Macau with sparse-binary side features (ChEMBL-like 170k compounds × 100k ECFP bits, 64 bits per compound, 1.5M
activities over 1000 targets, D=32) — exercises the beta path: sparse-binary SpMM pair + batched CG.

    python tools/bench_macau.py [scale]                                   one GPU
    torchrun --nproc-per-node N tools/bench_macau.py [scale]              N GPUs: rows sharded, the CG split by right-hand-side column

Prints one JSON line: sweeps/s, CG iterations, and the roofline of the (FᵀF + λI)·X product (two gather kernels): algorithmic gathered
bytes per application ÷ its device time against the L2 read bandwidth measured on the box by tools/l2_peak (the operand rows — beta
25.6 MB, F·beta 43.5 MB — are L2 hits; only the 87 MB of indices stream from HBM) and, for reference, against the HBM copy peak."""
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bdf_b200
from bdf_b200.multi import Comm
from tools.workloads import c3_macau

scale = float(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 1.0
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
D = 32
w = c3_macau(scale, D)
N, NT, NNZ, NUMF, BITS, ids, vals = w["N"], w["NT"], w["NNZ"], w["NUMF"], w["BITS"], w["ids"], w["vals"]
F = bdf_b200.SparseBinMatrix(w["rows"], w["cols"], N, NUMF)
comm = None
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = Comm(rank, world, local, D)
    stream = comm.stream
else:
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
eng = bdf_b200.Engine(D, device=local, rank=rank, world=world)
eng.set_stream(stream.cuda_stream)
eng.set_seed(20161017 + 2)
if comm is not None:
    from bdf_b200.shard import balanced_partition

    e1 = eng.add_entity_partitioned(N, balanced_partition(np.bincount(ids[:, 0] - 1, minlength=N), world, 2.0 * D))
    e2 = eng.add_entity_partitioned(NT, balanced_partition(np.bincount(ids[:, 1] - 1, minlength=NT), world, 2.0 * D))
else:
    e1, e2 = eng.add_entity(N), eng.add_entity(NT)
rel = eng.add_relation([e1, e2], ids, vals)
eng.set_relation_params(rel, 5.0, float(vals.mean()))
eng.set_features(e1, F)
if comm is not None:
    comm.connect(eng, [e1, e2], [e1])
eng.set_async(True)
mu, Lam = np.zeros(D), 5.0 * np.eye(D)
mu2, Lam2 = np.zeros(D), 5.0 * np.eye(D)
lb = 1.0
t_beta = []


def sweep():
    """One Gibbs iteration of src/macau.jl:96-140 for this model (entity 1 with features, entity 2 without)."""
    global mu, Lam, mu2, Lam2, lb
    eng.update_uhat(e1, mu)
    eng.sample_mode_uhat(e1, Lam, None)
    if comm is not None:
        comm.nw_stats(eng, e1, uhat=True)
    else:
        eng.nw_stats_uhat(e1)
    eng.nw_sample_async(e1, np.zeros(D), 2.0, np.eye(D) + eng.beta_gram(e1) * lb, float(D + NUMF))
    eng.sample_mode(e2, mu2, Lam2, None)
    if comm is not None:
        comm.nw_stats(eng, e2, uhat=False)
    else:
        eng.step_nw_stats(e2)
    eng.nw_sample_async(e2, np.zeros(D), 2.0, np.eye(D), float(D))
    mu, Lam = eng.nw_sample_fetch(e1)
    mu2, Lam2 = eng.nw_sample_fetch(e2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, iters = eng.sample_beta(e1, mu, Lam, lb, want_beta=False)
    if comm is not None:
        iters = np.asarray(comm.allreduce_scalars(iters))
    torch.cuda.synchronize()
    t_beta.append(time.perf_counter() - t0)
    lb, _ = eng.sample_lambda_beta(e1, Lam, 1e-3, 1.0)
    eng.advance_sweep()
    return iters


for _ in range(2):
    sweep()
if comm is not None:
    comm.barrier()
torch.cuda.synchronize()
t_beta.clear()
t0 = time.perf_counter()
K = 3
its = []
for _ in range(K):
    its.append(int(np.max(sweep())))
if comm is not None:
    comm.barrier()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / K
if rank == 0:
    ms_ata = eng.debug_ata_time(e1, 20) if world == 1 else None
    nnzF = N * BITS
    alg_bytes = 2 * nnzF * (4 + 8 * D) + (N + NUMF) * D * 8 * 2   # index + gathered operand bytes of both products + outputs
    try:
        peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    l2 = None
    try:
        l2 = json.loads(subprocess.run(["tools/l2_peak"], capture_output=True, text=True, timeout=60).stdout)["l2_read_gbs"]
    except Exception:
        pass
    line = {"config": f"C3 Macau {N}x{NT}, {NNZ} activities, F {N}x{NUMF} with {N*BITS} bits, D={D}", "n_gpus": world, "sweeps_per_s": 1 / dt,
            "ms_per_sweep": dt * 1e3, "ms_beta_draw": float(np.mean(t_beta)) * 1e3, "cg_iterations_max_per_sweep": its, "lambda_beta": lb,
            "beta_solve": "CG split by right-hand-side column over the ranks, solved columns stored into every replica (NVLink)" if world > 1 else "CG, all columns on one GPU"}
    if ms_ata is not None:
        line["ata_mul"] = {"ms": ms_ata, "algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / ms_ata / 1e6,
                           "bound": "L2 gather bandwidth (operand rows are L2 hits; 87 MB of indices per application stream from HBM)",
                           "peak_l2_read_gbs": l2, "frac_of_l2_peak": (alg_bytes / ms_ata / 1e6 / l2) if l2 else None,
                           "hbm_copy_peak_gbs": peak, "ratio_to_hbm_copy_peak": alg_bytes / ms_ata / 1e6 / peak}
    print(json.dumps(line))
eng.close()
if comm is not None:
    torch.distributed.destroy_process_group()
