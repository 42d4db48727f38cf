"""ctypes front-end for oracle/bdf_oracle.c plus a numpy/scipy twin of the row draw.

TEST INFRASTRUCTURE ONLY — see the header of bdf_oracle.c. Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs import this module.

The twin (`twin_*` functions) restates the same reference lines with numpy and calls the very
LAPACK routines Julia Base calls for `inv` (dgetrf+dgetri) and `chol` (dpotrf 'U') through
scipy.linalg.lapack; tests/test_oracle.py checks the C restatement against it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "bdf_oracle.c")
_SO = os.path.join(_HERE, "_build", "libbdf_oracle.so")

c_dp = C.POINTER(C.c_double)
c_i64p = C.POINTER(C.c_int64)
c_i32p = C.POINTER(C.c_int32)


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (-O3 -march=native -fopenmp). Returns the .so path."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-o", _SO, _SRC, "-lm"]
        )
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        try:
            _lib = C.CDLL(build())
        except (OSError, subprocess.CalledProcessError):
            # a prebuilt .so from another host (-march=native) may not load: rebuild once
            _lib = C.CDLL(build(force=True))
        _lib.orc_lambda_beta.restype = C.c_double
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def max_threads() -> int:
    return int(lib().orc_max_threads())


# ----------------------------------------------------------------------------------------------
# dense helpers
# ----------------------------------------------------------------------------------------------
def inv(A):
    """Julia inv(A): LU inverse. A is square; returns a new array (column-major semantics)."""
    A = np.asfortranarray(A, dtype=np.float64).copy(order="F")
    n = A.shape[0]
    info = lib().orc_inv(n, _dp(A))
    if info:
        raise np.linalg.LinAlgError(f"singular (info={info})")
    return A


def chol_upper(A):
    A = np.asfortranarray(A, dtype=np.float64)
    n = A.shape[0]
    R = np.zeros((n, n), order="F")
    info = lib().orc_chol_upper(n, _dp(A), _dp(R))
    if info:
        raise np.linalg.LinAlgError(f"not positive definite (info={info})")
    return R


# ----------------------------------------------------------------------------------------------
# observation index (IndexedDF / FastIDF)
# ----------------------------------------------------------------------------------------------
class FastIDF:
    """src/IndexedDF.jl:46-70. ids: nnz×K 1-based int64 (stored column-major), values: nnz float64."""

    def __init__(self, ids, values, dims):
        ids = np.asarray(ids, dtype=np.int64)
        self.nnz, self.K = ids.shape
        self.ids = np.asfortranarray(ids)
        self.values = _f64(values)
        self.dims = np.asarray(dims, dtype=np.int64)
        self.ptr = [np.zeros(int(d) + 1, dtype=np.int64) for d in self.dims]
        self.pos = [np.zeros(max(self.nnz, 1), dtype=np.int64) for _ in self.dims]
        ptrs = (c_i64p * self.K)(*[p.ctypes.data_as(c_i64p) for p in self.ptr])
        poss = (c_i64p * self.K)(*[p.ctypes.data_as(c_i64p) for p in self.pos])
        rc = lib().orc_idf_build(
            self.K, self.dims.ctypes.data_as(c_i64p), C.c_int64(self.nnz), self.ids.ctypes.data_as(c_i64p), ptrs, poss
        )
        if rc:
            raise ValueError("id out of range")

    def getData(self, mode: int, i: int):
        """mode, i are 1-based like Julia. Returns (ids[idx,:], values[idx])."""
        p = self.ptr[mode - 1]
        idx = self.pos[mode - 1][p[i - 1] : p[i]]
        return self.ids[idx, :], self.values[idx]

    def getCount(self, mode: int, i: int) -> int:
        p = self.ptr[mode - 1]
        return int(p[i] - p[i - 1])


# ----------------------------------------------------------------------------------------------
# half-sweep
# ----------------------------------------------------------------------------------------------
def sample_latent_all(idf: FastIDF, mode: int, U: list, alpha: float, mean_value: float, mu, Lambda, Z,
                      linear_values=None, nshards: int = 1):
    """sample_latent_all2! (src/sampling.jl:149-172) for 0-based `mode`; overwrites U[mode] (N×D C-order,
    i.e. Julia's D×N column-major). mu: (D,) or (N,D). Z: (N,D) injected standard normals."""
    D = U[mode].shape[1]
    for u in U:
        assert u.flags.c_contiguous and u.dtype == np.float64 and u.shape[1] == D
    mu = _f64(mu)
    mu_ld = 0 if mu.ndim == 1 else D
    Lambda = np.asfortranarray(Lambda, dtype=np.float64)
    Z = _f64(Z)
    Uptr = (c_dp * idf.K)(*[_dp(u) for u in U])
    lv = _dp(_f64(linear_values)) if linear_values is not None else None
    rc = lib().orc_sample_latent_all(
        D, idf.K, mode, idf.dims.ctypes.data_as(c_i64p), C.c_int64(idf.nnz), idf.ids.ctypes.data_as(c_i64p),
        _dp(idf.values), idf.ptr[mode].ctypes.data_as(c_i64p), idf.pos[mode].ctypes.data_as(c_i64p), Uptr,
        C.c_double(alpha), C.c_double(mean_value), lv, _dp(mu), C.c_int64(mu_ld), _dp(Lambda), _dp(Z), int(nshards),
    )
    if rc:
        raise np.linalg.LinAlgError("row draw failed (singular / not PD)")
    return U[mode]


class _RowData(C.Structure):
    _fields_ = [
        ("nother", C.c_int), ("U", C.POINTER(c_dp)), ("ids", C.POINTER(c_i64p)), ("vals", c_dp),
        ("offsets", c_dp), ("offset", C.c_double), ("alpha", C.c_double), ("nobs", C.c_int64),
    ]


def sample_row(D, rels, mu, Lambda, z):
    """General per-row draw (sample_user2, src/sampling.jl:266-289). rels: list of dicts with keys
    U (list of N_j×D arrays), ids (list of 1-based int64 arrays), vals, offset (float) or offsets (array), alpha."""
    keep = []
    arr = (_RowData * len(rels))()
    for r, rel in enumerate(rels):
        Us = [_f64(u) for u in rel["U"]]
        ids = [np.ascontiguousarray(i, dtype=np.int64) for i in rel["ids"]]
        vals = _f64(rel["vals"])
        offs = _f64(rel["offsets"]) if rel.get("offsets") is not None else None
        up = (c_dp * len(Us))(*[_dp(u) for u in Us])
        ip = (c_i64p * len(ids))(*[i.ctypes.data_as(c_i64p) for i in ids])
        keep += [Us, ids, vals, offs, up, ip]
        arr[r].nother = len(Us)
        arr[r].U = up
        arr[r].ids = ip
        arr[r].vals = _dp(vals)
        arr[r].offsets = _dp(offs) if offs is not None else None
        arr[r].offset = float(rel.get("offset", 0.0))
        arr[r].alpha = float(rel["alpha"])
        arr[r].nobs = len(vals)
    mu = _f64(mu)
    Lambda = np.asfortranarray(Lambda, dtype=np.float64)
    z = _f64(z)
    x = np.zeros(D)
    rc = lib().orc_sample_row(D, len(rels), arr, _dp(mu), _dp(Lambda), _dp(z), _dp(x), None)
    if rc:
        raise np.linalg.LinAlgError(f"row draw failed ({rc})")
    return x


# ----------------------------------------------------------------------------------------------
# Normal-Wishart
# ----------------------------------------------------------------------------------------------
def nw_stats(U, uhat=None):
    U = _f64(U)
    N, D = U.shape
    NU = np.zeros(D)
    NS = np.zeros((D, D), order="F")
    uh = _dp(_f64(uhat)) if uhat is not None else None
    lib().orc_nw_stats(D, C.c_int64(N), _dp(U), uh, _dp(NU), _dp(NS))
    return float(N), NU, NS


def cond_normal_wishart(N, NU, NS, mu0, beta0, Tinv, nu):
    D = len(NU)
    mu_N = np.zeros(D)
    T_N = np.zeros((D, D), order="F")
    beta_N = C.c_double()
    nu_N = C.c_double()
    rc = lib().orc_cond_normal_wishart(
        D, C.c_double(N), _dp(_f64(NU)), _dp(np.asfortranarray(NS, dtype=np.float64)), _dp(_f64(mu0)), C.c_double(beta0),
        _dp(np.asfortranarray(Tinv, dtype=np.float64)), C.c_double(nu), _dp(mu_N), C.byref(beta_N), _dp(T_N), C.byref(nu_N),
    )
    if rc:
        raise np.linalg.LinAlgError("T_N inverse failed")
    return mu_N, beta_N.value, T_N, nu_N.value


def nw_rand(mu_N, kappa, T, A, z):
    D = len(mu_N)
    mu = np.zeros(D)
    Lam = np.zeros((D, D), order="F")
    rc = lib().orc_nw_rand(
        D, _dp(_f64(mu_N)), C.c_double(kappa), _dp(np.asfortranarray(T, dtype=np.float64)),
        _dp(np.asfortranarray(A, dtype=np.float64)), _dp(_f64(z)), _dp(mu), _dp(Lam),
    )
    if rc:
        raise np.linalg.LinAlgError(f"nw_rand failed ({rc})")
    return mu, Lam


def bartlett_factor(rng: np.random.Generator, D: int, nu: float):
    """The random part of Distributions.jl's Wishart sampler [ext]: lower-triangular A with
    A[i,i] = sqrt(chi2(nu - i)) (0-based i), A[i>j] ~ N(0,1)."""
    A = np.zeros((D, D), order="F")
    for i in range(D):
        A[i, i] = np.sqrt(rng.chisquare(nu - i))
        for j in range(i):
            A[i, j] = rng.standard_normal()
    return A


def sample_beta_rel(F, values, udot, mean_value, alpha, lambda_beta, z1, z2):
    """sample_beta_rel — src/sampling.jl:322-337: res = values − udot − mean; aFt_y = α·F'(res + α^(-1/2)·randn(N)) + sqrt(λ)·randn(F);
    K = α·FF + λ·I; K \\ aFt_y (LU, as Julia's `\\`). z1, z2 are the injected randn(N), randn(F) (drawn in that order)."""
    F = np.asarray(F, dtype=np.float64)
    res = np.asarray(values, dtype=np.float64) - np.asarray(udot, dtype=np.float64) - mean_value
    aFt_y = alpha * (F.T @ (res + alpha ** -0.5 * np.asarray(z1))) + np.sqrt(lambda_beta) * np.asarray(z2)
    return solve_full(alpha * (F.T @ F), aFt_y.reshape(-1, 1), lambda_beta)[:, 0]


def sample_alpha(alpha_lambda0: float, alpha_nu0: float, err, chi2: float) -> float:
    """sample_alpha — src/sampling.jl:129-134: Λ = alpha_lambda0·eye(1); SW = inv(inv(Λ) + err'err);
    rand(Wishart(alpha_nu0 + n, SW))[1]. A 1×1 Wishart(ν, S) draw is S·chi2(ν) (Bartlett with a single diagonal entry
    [ext]); `chi2` is that injected variate."""
    err = np.asarray(err, dtype=np.float64)
    SW = 1.0 / (1.0 / alpha_lambda0 + float(err @ err))
    return SW * chi2


# ----------------------------------------------------------------------------------------------
# sparse binary operators, CG, β
# ----------------------------------------------------------------------------------------------
def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def sbm_mul(m, n, rows, cols, x):
    rows, cols, x = _i32(rows), _i32(cols), _f64(x)
    y = np.zeros(m)
    lib().orc_sbm_mul(C.c_int64(m), C.c_int64(n), C.c_int64(len(rows)), rows.ctypes.data_as(c_i32p), cols.ctypes.data_as(c_i32p), _dp(x), _dp(y))
    return y


def sbm_tmul(m, n, rows, cols, x):
    rows, cols, x = _i32(rows), _i32(cols), _f64(x)
    y = np.zeros(n)
    lib().orc_sbm_tmul(C.c_int64(m), C.c_int64(n), C.c_int64(len(rows)), rows.ctypes.data_as(c_i32p), cols.ctypes.data_as(c_i32p), _dp(x), _dp(y))
    return y


def csr_build(rows, cols):
    """SparseBinMatrixCSR(rows, cols) (src/sparsebin_csr.jl:22-37): returns m, n, row_ptr, col_ind (1-based Int32)."""
    rows, cols = _i32(rows), _i32(cols)
    m = C.c_int64()
    n = C.c_int64()
    row_ptr = np.zeros(int(rows.max()) + 1, dtype=np.int32)
    col_ind = np.zeros(len(rows), dtype=np.int32)
    lib().orc_csr_build(C.c_int64(len(rows)), rows.ctypes.data_as(c_i32p), cols.ctypes.data_as(c_i32p), C.byref(m), C.byref(n),
                        row_ptr.ctypes.data_as(c_i32p), col_ind.ctypes.data_as(c_i32p))
    return m.value, n.value, row_ptr, col_ind


def csr_mul(m, row_ptr, col_ind, x):
    x = _f64(x)
    y = np.zeros(m)
    lib().orc_csr_mul(C.c_int64(m), _i32(row_ptr).ctypes.data_as(c_i32p), _i32(col_ind).ctypes.data_as(c_i32p), _dp(x), _dp(y))
    return y


def csc_from_triplets(rows, cols, vals, m=None, n=None):
    """Julia's sparse(rows, cols, vals) (1-based triplets; duplicates are added, entries sorted by row inside a column) as the fields of a
    SparseMatrixCSC: (m, n, colptr, rowval, nzval), 1-based. [ext: Julia Base]"""
    rows, cols, vals = np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64), _f64(vals)
    m = int(rows.max()) if m is None else m
    n = int(cols.max()) if n is None else n
    order = np.lexsort((rows, cols))
    r, c, v = rows[order], cols[order], vals[order]
    keep = np.ones(len(r), dtype=bool)
    keep[1:] = (r[1:] != r[:-1]) | (c[1:] != c[:-1])
    starts = np.flatnonzero(keep)
    nz = np.add.reduceat(v, starts) if len(v) else v
    r, c = r[starts], c[starts]
    colptr = np.ones(n + 1, dtype=np.int64)
    np.add.at(colptr, c, 1)          # colptr[j+1] counts column j (1-based c → index c)
    colptr = np.cumsum(colptr) - np.arange(n + 1)
    return m, n, colptr, r, nz


def csc_mul(m, n, colptr, rowval, nzval, x):
    """A * x for a SparseMatrixCSC (Julia Base A_mul_B!: column by column, y[rowval[k]] += nzval[k]*x[j]) [ext]. This is what
    `At_mul_B(::SparseMatrixCSR, u)` runs on the stored transpose (src/parallel_csr.jl:41-42)."""
    y = np.zeros(m)
    for j in range(n):
        xj = x[j]
        for k in range(colptr[j] - 1, colptr[j + 1] - 1):
            y[rowval[k] - 1] += nzval[k] * xj
    return y


def csc_tmul(m, n, colptr, rowval, nzval, x):
    """At_mul_B(A, x) for a SparseMatrixCSC (Julia Base: y[j] = sum_k nzval[k]*x[rowval[k]] in stored order) [ext]. This is what
    `*(::SparseMatrixCSR, u)` runs on the stored transpose (src/parallel_csr.jl:43) — a row-wise CSR product."""
    y = np.zeros(n)
    for j in range(n):
        s = 0.0
        for k in range(colptr[j] - 1, colptr[j + 1] - 1):
            s += nzval[k] * x[rowval[k] - 1]
        y[j] = s
    return y


def sbm_ata_mul(m, n, rows, cols, x, lam):
    rows, cols, x = _i32(rows), _i32(cols), _f64(x)
    y = np.zeros(n)
    tmp = np.zeros(m)
    lib().orc_sbm_ata_mul(C.c_int64(m), C.c_int64(n), C.c_int64(len(rows)), rows.ctypes.data_as(c_i32p), cols.ctypes.data_as(c_i32p), _dp(x), C.c_double(lam), _dp(y), _dp(tmp))
    return y


def dense_ata_mul(F, x, lam):
    F = np.asfortranarray(F, dtype=np.float64)
    m, n = F.shape
    x = _f64(x)
    y = np.zeros(n)
    tmp = np.zeros(m)
    lib().orc_dense_ata_mul(C.c_int64(m), C.c_int64(n), _dp(F), _dp(x), C.c_double(lam), _dp(y), _dp(tmp))
    return y


def cg_ata(m, n, rows, cols, b, lam, tol=None, maxiter=None, F=None):
    """cg_AtA (src/parallel_cg.jl:63-94). Returns (x, operator applications)."""
    b = _f64(b)
    tol = n * np.finfo(float).eps if tol is None else tol
    maxiter = n if maxiter is None else maxiter
    x = np.zeros(n)
    if F is not None:
        F = np.asfortranarray(F, dtype=np.float64)
        its = lib().orc_cg_ata(C.c_int64(m), C.c_int64(n), C.c_int64(0), None, None, _dp(F), _dp(b), C.c_double(lam), C.c_double(tol), C.c_int64(maxiter), _dp(x))
    else:
        rows, cols = _i32(rows), _i32(cols)
        its = lib().orc_cg_ata(C.c_int64(m), C.c_int64(n), C.c_int64(len(rows)), rows.ctypes.data_as(c_i32p), cols.ctypes.data_as(c_i32p), None, _dp(b), C.c_double(lam), C.c_double(tol), C.c_int64(maxiter), _dp(x))
    return x, its


def solve_cg2(m, n, rows, cols, rhs, lam, tol=1e-6, maxiter=None, nthreads=1):
    """solve_cg2 (src/parallel_matrix.jl:488-507). rhs: n×ncol (any layout) → beta n×ncol Fortran-order, iters."""
    rhs = np.asfortranarray(rhs, dtype=np.float64)
    ncol = rhs.shape[1]
    maxiter = n if maxiter is None else maxiter
    rows, cols = _i32(rows), _i32(cols)
    beta = np.zeros((n, ncol), order="F")
    iters = np.zeros(ncol, dtype=np.int32)
    lib().orc_solve_cg2(C.c_int64(m), C.c_int64(n), C.c_int64(len(rows)), rows.ctypes.data_as(c_i32p), cols.ctypes.data_as(c_i32p), None, _dp(rhs), ncol,
                        C.c_double(lam), C.c_double(tol), C.c_int64(maxiter), _dp(beta), iters.ctypes.data_as(c_i32p), int(nthreads))
    return beta, iters


def solve_full(FF, rhs, lam):
    FF = np.asfortranarray(FF, dtype=np.float64)
    rhs = np.asfortranarray(rhs, dtype=np.float64)
    n = FF.shape[0]
    out = np.zeros(rhs.shape, order="F")
    rc = lib().orc_solve_full(C.c_int64(n), _dp(FF), _dp(rhs), rhs.shape[1], C.c_double(lam), _dp(out))
    if rc:
        raise np.linalg.LinAlgError("solve_full singular")
    return out


def color_noise(Lambda, E):
    """rand(MvNormal(0, inv(Λ)), ncols) with injected standard normals E (ncols×D C-order = D×ncols col-major)."""
    E = _f64(E)
    ncols, D = E.shape
    out = np.zeros_like(E)
    rc = lib().orc_color_noise(D, _dp(np.asfortranarray(Lambda, dtype=np.float64)), C.c_int64(ncols), _dp(E), _dp(out))
    if rc:
        raise np.linalg.LinAlgError("color_noise failed")
    return out


def beta_rhs_sbm(U, mu, N1, N2, rows, cols, numF, lambda_beta):
    """Ft_y of sample_beta (src/sampling.jl:300). U,N1: N×D C-order; N2: numF×D C-order. Returns numF×D Fortran-order."""
    U, N1, N2, mu = _f64(U), _f64(N1), _f64(N2), _f64(mu)
    N, D = U.shape
    rows, cols = _i32(rows), _i32(cols)
    out = np.zeros((numF, D), order="F")
    lib().orc_beta_rhs_sbm(D, C.c_int64(N), C.c_int64(numF), C.c_int64(len(rows)), rows.ctypes.data_as(c_i32p), cols.ctypes.data_as(c_i32p),
                           _dp(U), _dp(mu), _dp(N1), _dp(N2), C.c_double(lambda_beta), _dp(out))
    return out


def f_mul_beta_sbm(N, numF, rows, cols, beta):
    """uhat = (F·β)ᵀ (src/RelationData.jl:314-320): β numF×D Fortran-order → uhat N×D C-order."""
    beta = np.asfortranarray(beta, dtype=np.float64)
    D = beta.shape[1]
    rows, cols = _i32(rows), _i32(cols)
    uhat = np.zeros((N, D))
    lib().orc_f_mul_beta_sbm(D, C.c_int64(N), C.c_int64(numF), C.c_int64(len(rows)), rows.ctypes.data_as(c_i32p), cols.ctypes.data_as(c_i32p), _dp(beta), _dp(uhat))
    return uhat


def btb(beta):
    beta = np.asfortranarray(beta, dtype=np.float64)
    numF, D = beta.shape
    out = np.zeros((D, D), order="F")
    lib().orc_btb(D, C.c_int64(numF), _dp(beta), _dp(out))
    return out


def lambda_beta(BtB, Lambda, numF, nu, mu, g):
    D = BtB.shape[0]
    shape = C.c_double()
    v = lib().orc_lambda_beta(D, C.c_int64(numF), _dp(np.asfortranarray(BtB, dtype=np.float64)), _dp(np.asfortranarray(Lambda, dtype=np.float64)),
                              C.c_double(nu), C.c_double(mu), C.c_double(g), C.byref(shape))
    return float(v), shape.value


def pred(ids, U, mean_value):
    """udot + mean (src/sampling.jl:9-51). ids: ntest×K 1-based."""
    ids = np.asfortranarray(ids, dtype=np.int64)
    ntest, K = ids.shape
    D = U[0].shape[1]
    Us = [_f64(u) for u in U]
    Uptr = (c_dp * K)(*[_dp(u) for u in Us])
    out = np.zeros(ntest)
    lib().orc_pred(D, K, C.c_int64(ntest), ids.ctypes.data_as(c_i64p), Uptr, C.c_double(mean_value), _dp(out))
    return out


# ----------------------------------------------------------------------------------------------
# numpy/scipy twin of the row draw — same LAPACK entry points as Julia Base
# ----------------------------------------------------------------------------------------------
def twin_sample_user_basic(MM, rr, alpha, mu_u, Lambda_u, z):
    """src/sampling.jl:205-211 literally, MM is D×n (columns = partner factor vectors).
    covar = inv(Lambda_u + alpha*(MM*MM')) [getrf+getri]; mu = covar*(alpha*MM*rr + Lambda_u*mu_u);
    chol(Hermitian(covar))' * z + mu [potrf 'U']."""
    from scipy.linalg import lapack

    A = np.asfortranarray(Lambda_u + alpha * (MM @ MM.T))
    lu, piv, info = lapack.dgetrf(A)
    assert info == 0
    covar, info = lapack.dgetri(lu, piv)
    assert info == 0
    mu = covar @ (alpha * (MM @ rr) + Lambda_u @ mu_u)
    R, info = lapack.dpotrf(np.asfortranarray(covar), lower=0, clean=1)
    assert info == 0
    return R.T @ z + mu


def twin_sample_row_ul(MM, rr, alpha, mu_u, Lambda_u, z):
    """The algebraically identical inverse-free form the CUDA kernel uses (DESIGN.md §row draw):
    Λ* = W·Wᵀ with W upper-triangular ("UL" Cholesky = Cholesky of the index-reversed matrix),
    x = W⁻ᵀ(z + W⁻¹·rhs). Used by tests to quantify κ(Λ*)-amplified rounding between the two forms."""
    from scipy.linalg import solve_triangular

    D = len(mu_u)
    A = Lambda_u + alpha * (MM @ MM.T)
    rhs = alpha * (MM @ rr) + Lambda_u @ mu_u
    J = np.arange(D)[::-1]
    L = np.linalg.cholesky(A[np.ix_(J, J)])
    w = solve_triangular(L, rhs[J], lower=True)
    xr = solve_triangular(L.T, z[J] + w, lower=False)
    return xr[J]
