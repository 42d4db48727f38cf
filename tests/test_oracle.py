"""The CPU oracle against the reference's own known-answer fixtures and its scipy/LAPACK twin.

Reference tests restated here (paths relative to the reference checkout):
  test/sparsebin_csr.jl:4-23, test/parallel_matrix.jl:41-109, test/solver.jl:4-20,
  test/basic.jl:7-49, test/parallel_latent_basic.jl:4-30, test/heavy_copyto.jl:27-50.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as orc

# fixture of test/sparsebin_csr.jl:4-5 and test/parallel_matrix.jl:41-42
ROWS = np.concatenate([np.arange(1, 201), np.arange(151, 351)]).astype(np.int32)
COLS = np.concatenate([np.arange(151, 351), np.arange(1, 400, 2)]).astype(np.int32)


def approx(a, b, n=None):
    """Base.Test @test_approx_eq: |a-b| <= 1e4 * length * eps * max(|a|,|b|) (elementwise max norm form)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    n = a.size if n is None else n
    tol = 1e4 * n * np.finfo(float).eps
    assert np.max(np.abs(a - b)) <= tol * max(np.max(np.abs(a)), np.max(np.abs(b)), 1e-300)


def test_csr_ctor_known_answers():
    m, n, row_ptr, col_ind = orc.csr_build(ROWS, COLS)
    assert (m, n) == (350, 399)  # test/parallel_matrix.jl:47-48
    assert len(col_ind) == 400 and len(row_ptr) == 351
    # hand-derived from src/sparsebin_csr.jl:22-37
    assert list(row_ptr[0:5]) == [1, 2, 3, 4, 5]
    assert list(row_ptr[148:156]) == [149, 150, 151, 153, 155, 157, 159, 161]
    assert list(row_ptr[348:351]) == [399, 400, 401]
    for r, want in ((151, [301, 1]), (152, [302, 3]), (153, [303, 5])):
        got = col_ind[row_ptr[r - 1] - 1 : row_ptr[r] - 1]
        assert list(got) == want
    counts = np.diff(row_ptr.astype(np.int64))
    assert (counts == 1).sum() == 300 and (counts == 2).sum() == 50


def test_csr_ctor_empty_interior_rows_inherit_next_start():
    rows = np.array([1, 4, 4, 6], dtype=np.int32)
    cols = np.array([2, 1, 3, 3], dtype=np.int32)
    m, n, row_ptr, col_ind = orc.csr_build(rows, cols)
    assert (m, n) == (6, 3)
    assert list(row_ptr) == [1, 2, 2, 2, 4, 4, 5]
    assert list(col_ind) == [2, 1, 3, 3]


def test_spmv_csr_and_coo_match_sparse():
    rng = np.random.default_rng(0)
    m, n, row_ptr, col_ind = orc.csr_build(ROWS, COLS)
    x = rng.random(n)
    A = sp.coo_matrix((np.ones(len(ROWS)), (ROWS - 1, COLS - 1)), shape=(m, n)).tocsr()
    approx(orc.csr_mul(m, row_ptr, col_ind, x), A @ x)  # test/sparsebin_csr.jl:22
    approx(orc.sbm_mul(m, n, ROWS, COLS, x), A @ x)  # test/sparsebin_csr.jl:23
    x9 = rng.random(m)
    approx(orc.sbm_tmul(m, n, ROWS, COLS, x9), A.T @ x9)  # test/parallel_matrix.jl:71-76
    # AtA_mul_B!, λ=0.1: test/parallel_matrix.jl:78-92
    want = A.T @ (A @ x) + 0.1 * x
    approx(orc.sbm_ata_mul(m, n, ROWS, COLS, x, 0.1), want)
    approx(orc.dense_ata_mul(A.toarray(), x, 0.1), want)


def test_dense_ata_and_solve_full():
    rng = np.random.default_rng(1)
    A = rng.random((500, 20))
    x = rng.random(20)
    approx(orc.dense_ata_mul(A, x, 0.5), (A.T @ A + 0.5 * np.eye(20)) @ x)  # test/solver.jl:14-20
    X = rng.random((1000, 50))
    y = rng.random((50, 3))
    b2 = orc.solve_full(X.T @ X, y, 0.75)  # test/solver.jl:4-11
    approx(b2, np.linalg.solve(X.T @ X + 0.75 * np.eye(50), y))


def test_cg_matches_dense_solve():
    rng = np.random.default_rng(2)
    m, n = 350, 399
    A = sp.coo_matrix((np.ones(len(ROWS)), (ROWS - 1, COLS - 1)), shape=(m, n)).toarray()
    AA = A.T @ A
    x = rng.random(n)
    beta, its = orc.cg_ata(m, n, ROWS, COLS, x, 0.5)  # test/parallel_matrix.jl:107-109
    approx(beta, np.linalg.solve(AA + 0.5 * np.eye(n), x))
    assert 0 < its <= n
    beta2, _ = orc.cg_ata(m, n, ROWS, COLS, x, 0.75, tol=1e-6, maxiter=n)  # test/heavy_copyto.jl:36-44
    assert np.allclose(beta2, np.linalg.solve(AA + 0.75 * np.eye(n), x), rtol=0, atol=1e-5)
    rhs = rng.random((n, 3))
    Y, iters = orc.solve_cg2(m, n, ROWS, COLS, rhs, 0.5, tol=1e-6)  # test/heavy_copyto.jl:46-50
    assert np.allclose(Y, np.linalg.solve(AA + 0.5 * np.eye(n), rhs), rtol=0, atol=1e-5)
    assert iters.shape == (3,)
    # dense operator through the same CG
    F = rng.random((60, 12))
    b = rng.random(12)
    xd, _ = orc.cg_ata(60, 12, None, None, b, 0.3, maxiter=200, F=F)
    assert np.allclose(xd, np.linalg.solve(F.T @ F + 0.3 * np.eye(12), b), rtol=1e-9, atol=1e-12)


def test_indexed_df_known_answers():
    # test/basic.jl:7-31
    ids = np.array([[2, 1], [2, 3], [3, 4]])
    vals = np.array([0.0, -1.0, 0.5])
    X = orc.FastIDF(ids, vals, [4, 4])
    assert X.nnz == 3
    assert X.getData(1, 1)[0].shape == (0, 2) and X.getData(1, 1)[1].shape == (0,)
    assert X.getData(1, 4)[0].shape == (0, 2)
    i12, v12 = X.getData(1, 2)
    assert list(i12[:, 0]) == [2, 2] and list(i12[:, 1]) == [1, 3] and list(v12) == [0.0, -1.0]
    assert X.getData(2, 2)[0].shape == (0, 2)
    assert X.getCount(1, 2) == 2
    with pytest.raises(ValueError):
        orc.FastIDF(np.array([[5, 1]]), np.array([1.0]), [4, 4])
    # duplicates are legal and kept in table order (src/IndexedDF.jl:13-18)
    Xd = orc.FastIDF(np.array([[1, 2], [1, 2], [1, 1]]), np.array([1.0, 2.0, 3.0]), [2, 2])
    i, v = Xd.getData(1, 1)
    assert list(v) == [1.0, 2.0, 3.0] and list(i[:, 1]) == [2, 2, 1]


def test_inv_and_chol_against_lapack():
    from scipy.linalg import lapack

    rng = np.random.default_rng(3)
    for n in (1, 2, 5, 10, 32, 100):
        B = rng.standard_normal((n, n + 3))
        A = np.asfortranarray(B @ B.T + 0.5 * np.eye(n))
        lu, piv, info = lapack.dgetrf(A)
        want, info = lapack.dgetri(lu, piv)
        got = orc.inv(A)
        assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want)) * np.linalg.cond(A)
        R, info = lapack.dpotrf(A, lower=0, clean=1)
        assert np.max(np.abs(orc.chol_upper(A) - R)) <= 1e-13 * np.max(np.abs(R)) * np.linalg.cond(A)


def _latent_basic_fixture(seed=4):
    """test/parallel_latent_basic.jl:4-17 — 50×10 rank-2 data, D=5, α=5.0 (RelationData default, src/RelationData.jl:260)."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((50, 2))
    B = rng.standard_normal((10, 2))
    W = A @ B.T
    ii, jj = np.nonzero(np.ones_like(W))
    keep = rng.permutation(len(ii))[50:]  # assignToTest!(rel, 50)
    ids = np.stack([ii[keep] + 1, jj[keep] + 1], axis=1)
    vals = W[ii[keep], jj[keep]]
    idf = orc.FastIDF(ids, vals, [50, 10])
    D = 5
    U = [rng.standard_normal((50, D)) * 0.3, rng.standard_normal((10, D)) * 0.3]
    mu = rng.standard_normal(D) * 0.1
    G = rng.standard_normal((D, D))
    Lambda = G @ G.T + 2 * np.eye(D)
    return idf, U, mu, Lambda, vals.mean(), rng


def test_row_draw_two_paths_agree_and_match_lapack_twin():
    """test/parallel_latent_basic.jl:23-30: sample_user2 ≈ sample_user_basic for all 50 rows under one seed;
    plus: both ≈ the scipy twin that calls dgetrf/dgetri/dpotrf like Julia does."""
    idf, U, mu, Lambda, mean_value, rng = _latent_basic_fixture()
    D = 5
    Z = rng.standard_normal((50, D))
    Ub = [u.copy() for u in U]
    orc.sample_latent_all(idf, 0, Ub, 5.0, mean_value, mu, Lambda, Z)
    for i in range(50):
        ids, v = idf.getData(1, i + 1)
        s2 = orc.sample_row(D, [dict(U=[U[1]], ids=[ids[:, 1]], vals=v, offset=mean_value, alpha=5.0)], mu, Lambda, Z[i])
        approx(Ub[0][i], s2)
        MM = U[1][ids[:, 1] - 1].T
        tw = orc.twin_sample_user_basic(MM, v - mean_value, 5.0, mu, Lambda, Z[i])
        assert np.max(np.abs(Ub[0][i] - tw)) <= 1e-12 * max(1.0, np.max(np.abs(tw)))
        ul = orc.twin_sample_row_ul(MM, v - mean_value, 5.0, mu, Lambda, Z[i])
        assert np.max(np.abs(ul - tw)) <= 1e-12 * max(1.0, np.max(np.abs(tw)))


def test_half_sweep_is_independent_of_shard_count_and_handles_empty_rows():
    idf, U, mu, Lambda, mean_value, rng = _latent_basic_fixture(5)
    Z = rng.standard_normal((10, 5))
    outs = []
    for shards in (1, 2, 3, 8):
        Ub = [u.copy() for u in U]
        orc.sample_latent_all(idf, 1, Ub, 5.0, mean_value, mu, Lambda, Z, nshards=shards)
        outs.append(Ub[1])
    for o in outs[1:]:
        assert np.array_equal(o, outs[0])
    # a row without observations draws from the prior N(mu, inv(Lambda)) (MM is D×0)
    idf2 = orc.FastIDF(np.array([[1, 1], [3, 2]]), np.array([0.5, -0.5]), [3, 2])
    Ue = [np.zeros((3, 5)), rng.standard_normal((2, 5))]
    Ze = rng.standard_normal((3, 5))
    orc.sample_latent_all(idf2, 0, Ue, 2.0, 0.0, mu, Lambda, Ze)
    cov = np.linalg.inv(Lambda)
    want = np.linalg.cholesky(cov) @ Ze[1] + mu
    assert np.allclose(Ue[0][1], want, rtol=1e-12, atol=1e-12)


def test_tensor_row_draw_matches_twin():
    """sample_user_basic for tensors (src/sampling.jl:215-234): MM = Hadamard product of the other modes' columns."""
    rng = np.random.default_rng(6)
    dims = [7, 6, 3]
    D = 4
    ids = np.stack([rng.integers(1, d + 1, 60) for d in dims], axis=1)
    vals = rng.standard_normal(60)
    idf = orc.FastIDF(ids, vals, dims)
    U = [rng.standard_normal((d, D)) for d in dims]
    mu = rng.standard_normal(D)
    G = rng.standard_normal((D, D))
    Lambda = G @ G.T + np.eye(D)
    for mode in range(3):
        Z = rng.standard_normal((dims[mode], D))
        Ub = [u.copy() for u in U]
        orc.sample_latent_all(idf, mode, Ub, 1.5, 0.1, mu, Lambda, Z)
        others = [m for m in range(3) if m != mode]
        for i in range(dims[mode]):
            idr, v = idf.getData(mode + 1, i + 1)
            MM = np.ones((D, len(v)))
            for m in others:
                MM = MM * U[m][idr[:, m] - 1].T
            tw = orc.twin_sample_user_basic(MM, v - 0.1, 1.5, mu, Lambda, Z[i])
            assert np.max(np.abs(Ub[mode][i] - tw)) <= 1e-11 * max(1.0, np.max(np.abs(tw)))


def test_per_row_mu_matrix_and_linear_values():
    idf, U, mu, Lambda, mean_value, rng = _latent_basic_fixture(7)
    D = 5
    mu_mat = rng.standard_normal((50, D)) * 0.2
    lin = rng.standard_normal(idf.nnz) * 0.1 + mean_value
    Z = rng.standard_normal((50, D))
    Ub = [u.copy() for u in U]
    orc.sample_latent_all(idf, 0, Ub, 5.0, mean_value, mu_mat, Lambda, Z, linear_values=lin)
    for i in (0, 17, 49):
        p = idf.ptr[0]
        idx = idf.pos[0][p[i] : p[i + 1]]
        MM = U[1][idf.ids[idx, 1] - 1].T
        tw = orc.twin_sample_user_basic(MM, idf.values[idx] - lin[idx], 5.0, mu_mat[i], Lambda, Z[i])
        assert np.max(np.abs(Ub[0][i] - tw)) <= 1e-12 * max(1.0, np.max(np.abs(tw)))


def test_conditional_normal_wishart_and_draw():
    rng = np.random.default_rng(8)
    D, N = 6, 40
    U = rng.standard_normal((N, D))
    mu0 = np.zeros(D)
    n, NU, NS = orc.nw_stats(U)
    assert n == N
    approx(NU, U.sum(0))
    approx(NS, U.T @ U)
    mu_N, beta_N, T_N, nu_N = orc.cond_normal_wishart(n, NU, NS, mu0, 2.0, np.eye(D), D)
    # src/sampling.jl:116-127 literally
    want_mu = (2.0 * mu0 + U.sum(0)) / (2.0 + N)
    want_T = np.linalg.inv(np.eye(D) + U.T @ U + 2.0 * np.outer(mu0, mu0) - (2.0 + N) * np.outer(want_mu, want_mu))
    approx(mu_N, want_mu)
    approx(T_N, want_T)
    assert beta_N == 2.0 + N and nu_N == D + N
    A = orc.bartlett_factor(rng, D, nu_N)
    z = rng.standard_normal(D)
    mu, Lam = orc.nw_rand(mu_N, beta_N, T_N, A, z)
    Lc = np.linalg.cholesky(want_T)
    Zm = Lc @ A
    approx(Lam, Zm @ Zm.T)
    approx(mu, want_mu + np.linalg.cholesky(np.linalg.inv(Zm @ Zm.T) / beta_N) @ z)
    # uhat-subtracted statistics (src/macau.jl:124)
    uhat = rng.standard_normal((N, D))
    _, NU2, NS2 = orc.nw_stats(U, uhat)
    approx(NU2, (U - uhat).sum(0))
    approx(NS2, (U - uhat).T @ (U - uhat))


def test_beta_sampler_pieces():
    rng = np.random.default_rng(9)
    N, numF, D = 30, 12, 4
    dens = rng.random((N, numF)) < 0.3
    rows, cols = np.nonzero(dens)
    rows = (rows + 1).astype(np.int32)
    cols = (cols + 1).astype(np.int32)
    F = dens.astype(float)
    U = rng.standard_normal((N, D))
    mu = rng.standard_normal(D)
    G = rng.standard_normal((D, D))
    Lambda = G @ G.T + np.eye(D)
    E1 = rng.standard_normal((N, D))
    E2 = rng.standard_normal((numF, D))
    C = np.linalg.cholesky(np.linalg.inv(Lambda))
    N1 = orc.color_noise(Lambda, E1)
    approx(N1, E1 @ C.T)
    N2 = orc.color_noise(Lambda, E2)
    lb = 1.7
    rhs = orc.beta_rhs_sbm(U, mu, N1, N2, rows, cols, numF, lb)
    want = F.T @ ((U - mu) + N1) + np.sqrt(lb) * N2  # src/sampling.jl:300
    approx(rhs, want)
    beta = orc.solve_full(F.T @ F, rhs, lb)
    approx(beta, np.linalg.solve(F.T @ F + lb * np.eye(numF), want))
    beta_cg, its = orc.solve_cg2(N, numF, rows, cols, rhs, lb, tol=np.finfo(float).eps * numF)
    assert np.allclose(beta_cg, beta, rtol=1e-6, atol=1e-8)
    uhat = orc.f_mul_beta_sbm(N, numF, rows, cols, beta)
    approx(uhat, F @ beta)
    BtB = orc.btb(beta)
    approx(BtB, beta.T @ beta)
    lam, shape = orc.lambda_beta(BtB, Lambda, numF, 1e-3, 1.0, 0.9)
    nux = 1e-3 + numF * D
    mux = 1.0 * nux / (1e-3 + 1.0 * np.trace(beta.T @ beta @ Lambda))
    assert shape == nux / 2 and abs(lam - 0.9 * 2 * mux / nux) <= 1e-14 * lam


def test_pred_identity():
    # test/basic.jl:168-174 and test/tensor.jl:26-31
    rng = np.random.default_rng(10)
    U = [rng.standard_normal((5, 3)), rng.standard_normal((4, 3)), rng.standard_normal((2, 3))]
    ids = np.array([[4, 2, 1], [1, 1, 2]])
    out = orc.pred(ids, U, 0.25)
    approx(out[0], (U[0][3] * U[1][1] * U[2][0]).sum() + 0.25)
    out2 = orc.pred(ids[:, :2], U[:2], -1.0)
    approx(out2[1], U[0][0] @ U[1][0] - 1.0)


def test_restated_wishart_and_mvnormal_conventions_have_the_right_distribution():
    """No reference test pins the Distributions.jl / PDMats.jl streams, so the restated conventions — Bartlett factor with
    A[i,i] = sqrt(chi2(nu - i)), Z = chol_lower(T)·A, Lam = Z·Zᵀ; mu = mu_N + chol_lower(inv(Lam)/beta_N)·z — are at least pinned
    DISTRIBUTIONALLY against an independent implementation (scipy.stats): a small nu makes the marginals sensitive to the order of the
    chi-square degrees of freedom and to a lower/upper mix-up, which the moments at nu >> D are not."""
    from scipy import stats

    rng = np.random.default_rng(2024)
    D, nu, kappa, n = 4, 5.0, 3.0, 12000
    G = rng.standard_normal((D, D))
    T = G @ G.T / D + 0.3 * np.eye(D)
    T[0, 0] *= 6.0                                  # unequal scales: a permuted dof order would show in the diagonal marginals
    T = (T + T.T) / 2
    mu_N = rng.standard_normal(D)
    Lam = np.empty((n, D, D))
    q = np.empty(n)
    for s in range(n):
        A = orc.bartlett_factor(rng, D, nu)
        mu, L = orc.nw_rand(mu_N, kappa, T, A, rng.standard_normal(D))
        Lam[s] = L
        q[s] = kappa * (mu - mu_N) @ L @ (mu - mu_N)   # ~ chi2(D) whatever Lam was
    ref = stats.wishart(df=nu, scale=T).rvs(n, random_state=np.random.default_rng(7))
    # first moment: E[Lam] = nu·T, element-wise within 5 standard errors (Var = nu (T_ij² + T_ii T_jj))
    se = np.sqrt(nu * (T ** 2 + np.outer(np.diag(T), np.diag(T))) / n)
    assert np.all(np.abs(Lam.mean(0) - nu * T) <= 5 * se)
    # marginals against scipy's sampler: diagonal elements, one off-diagonal, log-determinant
    for pick in (lambda M: M[:, 0, 0], lambda M: M[:, D - 1, D - 1], lambda M: M[:, 1, 1], lambda M: M[:, 2, 0],
                 lambda M: np.linalg.slogdet(M)[1]):
        assert stats.ks_2samp(pick(Lam), pick(ref)).pvalue > 1e-4
    # Lam_ii / T_ii ~ chi2(nu) exactly, for every i (this is what fails if the dof run the wrong way along the diagonal)
    for i in range(D):
        assert stats.kstest(Lam[:, i, i] / T[i, i], stats.chi2(nu).cdf).pvalue > 1e-4
    assert stats.kstest(q, stats.chi2(D).cdf).pvalue > 1e-4
    # sample_alpha's 1×1 Wishart and sample_lambda_beta's Gamma are plain scalings of the injected chi-square / gamma variate
    assert abs(orc.sample_alpha(2.0, 1.0, np.array([1.0, 2.0]), 3.0) - 3.0 / (0.5 + 5.0)) <= 1e-15


def test_sparse_csr_wrapper_known_answers():
    """test/sparse_csr.jl:4-34 — SparseMatrixCSR (src/parallel_csr.jl:36-54) stores the CSC of the transpose; `Xr*y` is the row-wise product
    on it, `At_mul_B(Xr, y)` the column-wise one; both equal the products of the original matrix, as does the 4-entry triplet fixture."""
    import scipy.sparse as sp

    rng = np.random.default_rng(12)
    X = sp.random(50, 100, 0.1, random_state=3, format="coo")             # sprand(50, 100, 0.1)
    r, c, v = X.row + 1, X.col + 1, X.data
    mt, nt, colptr, rowval, nzval = orc.csc_from_triplets(c, r, v, m=100, n=50)   # sparse_csr(X) = SparseMatrixCSR(X')
    assert (nt, mt) == X.shape                                             # size(Xr) == size(X): (n, m) of the stored transpose
    y1, y2 = rng.random(100), rng.random(50)
    Xd = X.toarray()
    approx(orc.csc_tmul(mt, nt, colptr, rowval, nzval, y1), Xd @ y1)       # Xr * y1
    approx(orc.csc_mul(mt, nt, colptr, rowval, nzval, y2), Xd.T @ y2)      # At_mul_B(Xr, y2) == Ac_mul_B(Xr, y2)
    # the products are also bit-identical to scipy's CSR / CSC kernels, which accumulate in the same (stored) order — the reason the
    # GPU test of general sparse features can compare against scipy bit for bit
    Xcsc = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(mt, nt))
    assert np.array_equal(orc.csc_mul(mt, nt, colptr, rowval, nzval, y2), Xcsc @ y2)
    assert np.array_equal(orc.csc_tmul(mt, nt, colptr, rowval, nzval, y1), sp.csr_matrix(Xcsc.T) @ y1)
    # At_mul_B(Xr, Xr) == At_mul_B(X, X)
    approx((Xcsc @ Xcsc.T).toarray(), Xd.T @ Xd)
    # rows = [1, 2, 2, 4]; cols = [2, 1, 3, 3]; vals = [0.1, 0.2, 0.15, 0.3]: sparse(rows, cols, vals) * z == sparse_csr(rows, cols, vals) * z
    rows, cols, vals = [1, 2, 2, 4], [2, 1, 3, 3], [0.1, 0.2, 0.15, 0.3]
    z = rng.random(3)
    m1, n1, cp1, rv1, nz1 = orc.csc_from_triplets(rows, cols, vals)        # sparse(rows, cols, vals): 4 x 3
    m2, n2, cp2, rv2, nz2 = orc.csc_from_triplets(cols, rows, vals)        # sparse_csr: SparseMatrixCSR(sparse(cols, rows, vals)): 3 x 4 stored
    assert (m1, n1, m2, n2) == (4, 3, 3, 4)
    w1 = orc.csc_mul(m1, n1, cp1, rv1, nz1, z)
    w2 = orc.csc_tmul(m2, n2, cp2, rv2, nz2, z)
    approx(w1, w2)
    approx(w1, np.array([0.1 * z[1], 0.2 * z[0] + 0.15 * z[2], 0.0, 0.3 * z[2]]))
