"""The reference's on-disk formats (src/data_reading.jl) — little-endian, Int64 headers, column-major payloads — so that
`output=` dumps written by this package load in Julia and feature files written by Julia load here. Host-side I/O only."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def write_binary_matrix(filename: str, X):
    """src/data_reading.jl:93-99: Int64 nrows, Int64 ncols, then X column-major in its own element type. X is given as the
    Julia matrix would be indexed, i.e. shape (nrows, ncols)."""
    X = np.asarray(X)
    with open(filename, "wb") as fh:
        np.array([X.shape[0], X.shape[1]], dtype="<i8").tofile(fh)
        np.asfortranarray(X).T.tofile(fh)  # column-major bytes


def _read_binary(filename: str, dtype):
    with open(filename, "rb") as fh:
        nrows, ncols = np.fromfile(fh, dtype="<i8", count=2)
        return np.fromfile(fh, dtype=dtype, count=int(nrows * ncols)).reshape((int(nrows), int(ncols)), order="F")


def read_binary_int32(filename: str):
    """src/data_reading.jl:53-59."""
    return _read_binary(filename, "<i4")


def read_binary_float32(filename: str):
    """src/data_reading.jl:61-67."""
    return _read_binary(filename, "<f4")


def write_sparse_float32(filename: str, rows, cols=None, values=None):
    """src/data_reading.jl:101-120: Int64 nnz, Int32 rows, Int32 cols (1-based), Float32 values. `rows` may be a scipy sparse matrix."""
    if cols is None:
        coo = rows.tocsc().tocoo()  # findnz order of a SparseMatrixCSC: column-major
        rows, cols, values = coo.row + 1, coo.col + 1, coo.data
    with open(filename, "wb") as fh:
        np.array([len(rows)], dtype="<i8").tofile(fh)
        np.asarray(rows, dtype="<i4").tofile(fh)
        np.asarray(cols, dtype="<i4").tofile(fh)
        np.asarray(values, dtype="<f4").tofile(fh)


def read_sparse_float32(filename: str):
    """src/data_reading.jl:69-77 → (rows, cols, vals), 1-based Int32 / Float32."""
    with open(filename, "rb") as fh:
        nnz = int(np.fromfile(fh, dtype="<i8", count=1)[0])
        rows = np.fromfile(fh, dtype="<i4", count=nnz)
        cols = np.fromfile(fh, dtype="<i4", count=nnz)
        vals = np.fromfile(fh, dtype="<f4", count=nnz)
        return rows, cols, vals


def write_sparse_binary_matrix(filename: str, X):
    """src/data_reading.jl:122-132: Int64 nrows, ncols, nnz; Int32 rows, cols (1-based) of the non-zeros, column-major order."""
    coo = sp.csc_matrix(X).tocoo()
    with open(filename, "wb") as fh:
        np.array([X.shape[0], X.shape[1], coo.nnz], dtype="<i8").tofile(fh)
        (coo.row + 1).astype("<i4").tofile(fh)
        (coo.col + 1).astype("<i4").tofile(fh)


def read_sparse_binary_matrix(filename: str):
    """src/data_reading.jl:134-143 → the feature matrix as a SparseBinMatrix (the COO lists the device path ingests directly)."""
    from .relation_data import SparseBinMatrix

    with open(filename, "rb") as fh:
        nrows, ncols, nnz = (int(v) for v in np.fromfile(fh, dtype="<i8", count=3))
        rows = np.fromfile(fh, dtype="<i4", count=nnz)
        cols = np.fromfile(fh, dtype="<i4", count=nnz)
    return SparseBinMatrix(rows, cols, nrows, ncols)


def write_sparse_float64(filename: str, X):
    """src/data_reading.jl:195-206."""
    coo = sp.csc_matrix(X).tocoo()
    with open(filename, "wb") as fh:
        np.array([X.shape[0], X.shape[1], coo.nnz], dtype="<i8").tofile(fh)
        (coo.row + 1).astype("<i4").tofile(fh)
        (coo.col + 1).astype("<i4").tofile(fh)
        coo.data.astype("<f8").tofile(fh)


def read_sparse_float64(filename: str):
    """src/data_reading.jl:208-218 → scipy CSC matrix."""
    with open(filename, "rb") as fh:
        nrow, ncol, nnz = (int(v) for v in np.fromfile(fh, dtype="<i8", count=3))
        rows = np.fromfile(fh, dtype="<i4", count=nnz)
        cols = np.fromfile(fh, dtype="<i4", count=nnz)
        vals = np.fromfile(fh, dtype="<f8", count=nnz)
    return sp.csc_matrix((vals, (rows - 1, cols - 1)), shape=(nrow, ncol))
