"""Generates tests/golden/movielens_1m.npz from the reference's own data file (BASELINE.json configs[0], "C1").

    python tests/golden/make_movielens_fixture.py            # needs /root/reference (this container only)

/root/reference/data/movielens_1m.mat holds X (6040×3952 ratings, 1 000 209 non-zeros with values 1..5; 246 item columns are
empty), Fu (6040×29) and Fv (3952×18), 0/1 side-information matrices (docs/index.md:34-60). The GPU box has no /root/reference,
so the three matrices are committed here in a compact lossless form: the non-zeros in Julia's `findnz` order of a
SparseMatrixCSC (column-major — the order `Relation(data::SparseMatrixCSC, …)` puts them into its table,
src/RelationData.jl:165-169), 0-based uint16 coordinates, uint8 values. `tests/movielens.py` turns them back into scipy
matrices."""
import os
import sys

import numpy as np
import scipy.io

SRC = "/root/reference/data/movielens_1m.mat"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "movielens_1m.npz")


def coo_colmajor(M):
    c = M.tocsc()
    c.sort_indices()
    c = c.tocoo()
    return c.row.astype(np.uint16), c.col.astype(np.uint16), c.data


def main():
    if not os.path.exists(SRC):
        sys.exit(f"{SRC} not found (the fixture is generated in the build container only)")
    d = scipy.io.loadmat(SRC)
    out = {}
    for key in ("X", "Fu", "Fv"):
        r, c, v = coo_colmajor(d[key])
        assert np.array_equal(v, np.round(v)) and v.min() >= 1 and v.max() <= 255
        out[f"{key}_row"], out[f"{key}_col"], out[f"{key}_val"] = r, c, v.astype(np.uint8)
        out[f"{key}_shape"] = np.asarray(d[key].shape, dtype=np.int64)
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT), "bytes;", {k: tuple(out[f"{k}_shape"]) for k in ("X", "Fu", "Fv")}, "nnz", len(out["X_val"]))


if __name__ == "__main__":
    main()
