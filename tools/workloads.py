"""Seeded generators of the synthetic workloads of BASELINE.json (configs C3 and C4; C2 lives in bench.py, C5 in tools/bench_c5.py),
shared by the bench tools and the scale-parity tests. SURVEY §8d: seed 20161017 + config index."""
import numpy as np

SEED = 20161017


def zipf_feature_bits(N, numF, bits, rng, s=1.1):
    """Sparse-binary side features of C3: exactly `bits` distinct set bits per row, bit popularity ~ Zipf(s). Weighted draws with
    replacement (inverse CDF), de-duplicated per row, rows that come out short are topped up with uniform bits. Returns 1-based
    Int32 COO lists (rows, cols) in row order, as the reference's SparseBinMatrix holds them (src/parallel_matrix.jl:9-24)."""
    w = 1.0 / np.arange(1, numF + 1) ** s
    cdf = np.cumsum(w / w.sum())
    draw = np.minimum(np.searchsorted(cdf, rng.random((N, 2 * bits))), numF - 1).astype(np.int32)
    draw.sort(axis=1)
    dup = np.zeros(draw.shape, dtype=bool)
    dup[:, 1:] = draw[:, 1:] == draw[:, :-1]
    key = rng.random(draw.shape)
    key[dup] = 2.0                                  # duplicates go last
    order = np.argsort(key, axis=1, kind="stable")[:, :bits]
    cols = np.take_along_axis(draw, order, axis=1)
    short = np.flatnonzero((~dup).sum(axis=1) < bits)
    for i in short:                                 # few rows (those whose draws hit the popular bits many times)
        u = np.unique(draw[i])
        extra = np.setdiff1d(rng.permutation(numF)[: 4 * bits], u)[: bits - len(u)]
        cols[i] = np.concatenate([u, extra]).astype(np.int32)
    rows = np.repeat(np.arange(1, N + 1, dtype=np.int32), bits)
    return rows, (cols.ravel() + 1).astype(np.int32)


def c3_macau(scale=1.0, D=32):
    """C3: 170k compounds × 1000 targets, 1.5M activities; F 170k × 100k bits, 64 per compound."""
    N, NT, NNZ, NUMF, BITS = int(170000 * scale), 1000, int(1500000 * scale), int(100000 * scale), 64
    rng = np.random.default_rng(SEED + 2)
    ids = np.stack([rng.integers(1, N + 1, NNZ), rng.integers(1, NT + 1, NNZ)], axis=1)
    vals = rng.standard_normal(NNZ)
    rows, cols = zipf_feature_bits(N, NUMF, BITS, rng)
    return {"N": N, "NT": NT, "NNZ": NNZ, "NUMF": NUMF, "BITS": BITS, "ids": ids, "vals": vals, "rows": rows, "cols": cols}


def c4_tensor(scale=1.0, dims=(20000, 5000, 200)):
    """C4: 3-mode tensor 20k × 5k × 200, 50M observed entries, skewed marginals idx = floor(N·u^2.5)."""
    NNZ = int(50_000_000 * scale)
    rng = np.random.default_rng(SEED + 3)
    ids = np.empty((NNZ, 3), dtype=np.int64, order="F")
    for m, d in enumerate(dims):
        ids[:, m] = np.minimum((d * rng.random(NNZ) ** 2.5).astype(np.int64), d - 1) + 1
    vals = rng.standard_normal(NNZ)
    return {"dims": list(dims), "NNZ": NNZ, "ids": ids, "vals": vals}
