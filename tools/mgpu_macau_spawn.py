"""python tools/mgpu_macau_spawn.py — `bdf_b200.macau(data, devices=[0, 1])` called from ONE ordinary process (no torchrun): the caller becomes
rank 0 and spawns a worker process for the second GPU (multi.macau_multi). Must reproduce the one-GPU run on the same seed."""
import sys

import numpy as np

sys.path.insert(0, ".")


def problem():
    import bdf_b200
    from bdf_b200.relation_data import assignToTest

    rng = np.random.default_rng(8)
    N, M, D0, nnz = 2500, 700, 3, 90000
    A, B = rng.standard_normal((N, D0)), rng.standard_normal((M, D0))
    ids = np.stack([np.minimum((N * rng.random(nnz) ** 2.0).astype(np.int64), N - 1) + 1, rng.integers(1, M + 1, nnz)], 1)
    vals = np.einsum("ij,ij->i", A[ids[:, 0] - 1], B[ids[:, 1] - 1]) + 0.3 * rng.standard_normal(nnz)
    rd = bdf_b200.RelationData((ids, vals, [N, M]), alpha=5.0, class_cut=0.0)
    assignToTest(rd.relations[0], 6000, np.random.default_rng(4))
    return rd


def count_calls(data):
    return float(np.linalg.norm(data.entities[0].model.sample))


if __name__ == "__main__":
    import bdf_b200

    kw = dict(num_latent=32, burnin=5, psamples=5, verbose=False, seed=21, clamp=[-6.0, 6.0])
    multi = bdf_b200.macau(problem(), devices=[0, 1], f=count_calls, **kw)     # the callback runs on rank 0 only and sees the live model
    single = bdf_b200.macau(problem(), device=0, f=count_calls, **kw)
    d_rmse = abs(multi["RMSE"] - single["RMSE"])
    d_pred = float(np.max(np.abs(multi["predictions"]["pred"] - single["predictions"]["pred"])))
    print(f"spawned 2-GPU run: RMSE {multi['RMSE']:.6f} vs one GPU {single['RMSE']:.6f}; max |pred diff| {d_pred:.2e}; f_output {multi['f_output'][-1]:.6f} vs {single['f_output'][-1]:.6f}")
    assert d_rmse < 1e-6 and d_pred < 1e-4 and len(multi["f_output"]) == 5
    assert abs(multi["f_output"][-1] - single["f_output"][-1]) < 1e-6 * single["f_output"][-1]
    print("MGPU MACAU SPAWN OK")
