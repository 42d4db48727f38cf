"""Generates tests/golden/*.npz — committed golden vectors for the row draw, made with the scipy/LAPACK twin
(oracle.twin_sample_user_basic: dgetrf+dgetri+dpotrf 'U', the routines Julia Base calls for `inv` and `chol`), i.e.
independently of both the C oracle and the CUDA kernel. The Julia reference itself cannot run in this image.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import oracle as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make(name, dims, nnz, D, alpha, seed):
    rng = np.random.default_rng(seed)
    K = len(dims)
    ids = np.stack([rng.integers(1, d + 1, nnz) for d in dims], axis=1).astype(np.int64)
    ids[ids[:, 0] == 2, 0] = 3  # row 2 of mode 1 has no observations → prior draw
    vals = rng.standard_normal(nnz)
    U = [rng.standard_normal((d, D)) * 0.4 for d in dims]
    G = rng.standard_normal((D, D)) * 0.3
    Lambda = G @ G.T + 2.0 * np.eye(D)
    mu = rng.standard_normal(D) * 0.2
    mean = float(vals.mean())
    idf = orc.FastIDF(ids, vals, dims)
    out = {"ids": ids, "vals": vals, "Lambda": Lambda, "mu": mu, "alpha": alpha, "mean": mean, "dims": np.array(dims)}
    for m in range(K):
        out[f"U{m}"] = U[m]
    for mode in range(K):
        Z = rng.standard_normal((dims[mode], D))
        X = np.zeros((dims[mode], D))
        for i in range(dims[mode]):
            idr, v = idf.getData(mode + 1, i + 1)
            MM = np.ones((D, len(v)))
            for m in range(K):
                if m != mode:
                    MM = MM * U[m][idr[:, m] - 1].T
            X[i] = orc.twin_sample_user_basic(MM, v - mean, alpha, mu, Lambda, Z[i])
        out[f"Z{mode}"] = Z
        out[f"X{mode}"] = X  # half-sweep of `mode` from the INITIAL factors (no Gauss-Seidel chaining)
    np.savez_compressed(os.path.join(HERE, name), **out)


if __name__ == "__main__":
    make("row_draw_matrix_d5.npz", [50, 10], 450, 5, 5.0, 1)       # test/parallel_latent_basic.jl shape
    make("row_draw_matrix_d32.npz", [60, 40], 1500, 32, 2.0, 2)
    make("row_draw_matrix_d100.npz", [30, 25], 900, 100, 1.5, 3)
    make("row_draw_tensor_d30.npz", [15, 12, 5], 800, 30, 2.0, 4)  # Khatri-Rao gather
    print("golden vectors written")
