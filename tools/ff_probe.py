"""solve_full at the reference's largest direct-solve size (numF = compute_ff_size = 6500): the blocked Cholesky of csrc/dense_spd.cuh
against numpy's LAPACK solve, and its time per call. Usage: python tools/ff_probe.py [numF] [D]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bdf_b200  # noqa: E402

numF = int(sys.argv[1]) if len(sys.argv) > 1 else 6500
D = int(sys.argv[2]) if len(sys.argv) > 2 else 32
rng = np.random.default_rng(3)
N = 20000
nnz_f = 40 * N
rows = rng.integers(1, N + 1, nnz_f).astype(np.int32)
cols = rng.integers(1, numF + 1, nnz_f).astype(np.int32)
key = np.unique(rows.astype(np.int64) * (numF + 1) + cols)
rows, cols = (key // (numF + 1)).astype(np.int32), (key % (numF + 1)).astype(np.int32)
eng = bdf_b200.Engine(D)
e1, e2 = eng.add_entity(N), eng.add_entity(5)
ids = np.stack([rng.integers(1, N + 1, 1000), rng.integers(1, 6, 1000)], axis=1)
eng.add_relation([e1, e2], ids, rng.standard_normal(1000))
eng.set_features(e1, bdf_b200.SparseBinMatrix(rows, cols, N, numF))
t = time.perf_counter()
FF = eng.compute_ff(e1, want=True)
t_ff = time.perf_counter() - t
rhs = rng.standard_normal((numF, D))
x = eng.solve_full(e1, rhs, 0.5)
ts = []
for _ in range(3):
    t = time.perf_counter()
    x = eng.solve_full(e1, rhs, 0.5)
    ts.append(time.perf_counter() - t)
t = time.perf_counter()
ref = np.linalg.solve(FF + 0.5 * np.eye(numF), rhs)
t_np = time.perf_counter() - t
err = float(np.max(np.abs(x - ref)) / np.max(np.abs(ref)))
print({"numF": numF, "D": D, "rel_err_vs_lapack": err, "solve_full_s_incl_copies": min(ts), "compute_ff_s": t_ff, "numpy_solve_s": t_np,
       "chol_gflop": numF ** 3 / 3e9})
assert err <= 1e-10
