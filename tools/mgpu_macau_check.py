"""torchrun --nproc-per-node 2 tools/mgpu_macau_check.py — `bdf_b200.macau(..., devices=[0, 1])` (every process of the group calls it,
rank r drives devices[r]) must reproduce the one-GPU `macau(...)` on the same seed: BPMF with a test set, and Macau with sparse-binary
side features (replicated beta path per rank). Philox noise is keyed by global row ids, so results agree to rounding."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import bdf_b200
from bdf_b200.relation_data import assignToTest

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def problem(with_features):
    rng = np.random.default_rng(3)
    N, M, D0, nnz = 3001, 803, 3, 120000
    A, B = rng.standard_normal((N, D0)), rng.standard_normal((M, D0))
    ids = np.stack([np.minimum((N * rng.random(nnz) ** 2.0).astype(np.int64), N - 1) + 1, rng.integers(1, M + 1, nnz)], 1)
    vals = np.einsum("ij,ij->i", A[ids[:, 0] - 1], B[ids[:, 1] - 1]) + 0.3 * rng.standard_normal(nnz)
    F = None
    if with_features:
        r, c = np.nonzero(rng.random((N, 40)) < 0.1)
        F = bdf_b200.SparseBinMatrix((r + 1).astype(np.int32), (c + 1).astype(np.int32), N, 40)
    rd = bdf_b200.RelationData((ids, vals, [N, M]), feat1=F, alpha=5.0, class_cut=0.0)
    assignToTest(rd.relations[0], 9000, np.random.default_rng(4))
    return rd


for feat in (False, True):
    for D in (16, 100):
        kw = dict(num_latent=D, burnin=6, psamples=6, verbose=False, seed=11, clamp=[-6.0, 6.0], compute_ff_size=0, full_prediction=(D == 16))
        multi = bdf_b200.macau(problem(feat), devices=list(range(world)), **kw)
        dist.barrier()
        if rank == 0:
            single = bdf_b200.macau(problem(feat), device=0, **kw)
            d_rmse = abs(multi["RMSE"] - single["RMSE"])
            d_pred = float(np.max(np.abs(multi["predictions"]["pred"] - single["predictions"]["pred"])))
            print(f"features={feat} D={D}: RMSE {multi['RMSE']:.6f} vs {single['RMSE']:.6f}; max |pred diff| {d_pred:.2e}; ROC {multi['ROC']:.4f} vs {single['ROC']:.4f}", flush=True)
            # the all-reduce adds the ranks' partial statistics in a different order than one GPU does: rounding-level differences in
            # (mu, Lambda) that 12 Gibbs sweeps carry along but do not blow up
            assert d_rmse < 1e-6 and d_pred < 1e-4, (d_rmse, d_pred)
            assert abs(multi["ROC"] - single["ROC"]) < 1e-6 and abs(multi["accuracy"] - single["accuracy"]) < 1e-3
            if "predictions_full" in single:   # pred_all from the slot-ordered replicas of a sharded run
                assert float(np.max(np.abs(multi["predictions_full"] - single["predictions_full"]))) < 1e-4
        dist.barrier()
if rank == 0:
    print("MGPU MACAU OK")
dist.destroy_process_group()
