"""GPU parity of the Macau link-matrix path (sparse-binary products, batched CG, beta / lambda_beta samplers, uhat)
against the CPU oracle, through the C ABI. Index structures and the 0/1 matvecs are bit-exact; floating-point results
within 1e-10 relative (looser only where the reference's own CG stopping rule leaves more slack, stated per test)."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

ROWS = np.concatenate([np.arange(1, 201), np.arange(151, 351)]).astype(np.int32)   # test/sparsebin_csr.jl:4-5
COLS = np.concatenate([np.arange(151, 351), np.arange(1, 400, 2)]).astype(np.int32)


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def engine_with_features(D, rows, cols, m, n, n2=7, seed=0):
    import bdf_b200

    rng = np.random.default_rng(seed)
    eng = bdf_b200.Engine(D)
    e1, e2 = eng.add_entity(m), eng.add_entity(n2)
    nnz = 5 * m
    ids = np.stack([rng.integers(1, m + 1, nnz), rng.integers(1, n2 + 1, nnz)], axis=1)
    vals = rng.standard_normal(nnz)
    rel = eng.add_relation([e1, e2], ids, vals)
    eng.set_features(e1, bdf_b200.SparseBinMatrix(rows, cols, m, n))
    return eng, e1, e2, rel, (ids, vals), rng


def test_csr_and_csc_structures_are_bit_exact():
    eng, e1, *_ = engine_with_features(8, ROWS, COLS, 350, 399)
    m, n, row_ptr, col_ind = orc.csr_build(ROWS, COLS)
    ptr, ind = eng.debug_features_csr(e1, False, len(ROWS))
    assert np.array_equal(ptr, row_ptr) and np.array_equal(ind, col_ind)
    assert list(ptr[148:156]) == [149, 150, 151, 153, 155, 157, 159, 161]  # SURVEY §8c known answers
    mt, nt, col_ptr, row_ind = orc.csr_build(COLS, ROWS)  # CSR of Fᵀ by the same constructor
    ptr_t, ind_t = eng.debug_features_csr(e1, True, len(ROWS))
    assert np.array_equal(ptr_t, col_ptr) and np.array_equal(ind_t, row_ind)
    eng.close()


@pytest.mark.parametrize("D", [1, 8, 10, 32, 100])
def test_spmm_is_bit_exact_with_the_reference_summation_order(D):
    rng = np.random.default_rng(D)
    m, n, nnz = 300, 257, 4000
    rows = rng.integers(1, m + 1, nnz).astype(np.int32)
    cols = rng.integers(1, n + 1, nnz).astype(np.int32)
    rows[:3] = [m, m, 1]
    cols[:3] = [n, n, 1]  # duplicates count twice; last row/col populated
    eng, e1, *_ = engine_with_features(D, rows, cols, m, n, seed=D)
    X = rng.standard_normal((n, D))
    Y = eng.spmm(e1, X)
    Xt = rng.standard_normal((m, D))
    Yt = eng.spmm(e1, Xt, transpose=True)
    for d in range(D):
        assert np.array_equal(Y[:, d], orc.sbm_mul(m, n, rows, cols, X[:, d]))       # A_mul_B!, list order
        assert np.array_equal(Yt[:, d], orc.sbm_tmul(m, n, rows, cols, Xt[:, d]))    # At_mul_B!
    x = rng.standard_normal(n)
    assert rel_err(eng.ata_mul(e1, x, 0.1), orc.sbm_ata_mul(m, n, rows, cols, x, 0.1)) <= 1e-15
    eng.close()


def test_reference_fixture_products_and_cg():
    """test/parallel_matrix.jl:41-109 and test/heavy_copyto.jl:27-50 on the device."""
    import scipy.sparse as sp

    D = 3
    eng, e1, *_ = engine_with_features(D, ROWS, COLS, 350, 399)
    rng = np.random.default_rng(1)
    A = sp.coo_matrix((np.ones(len(ROWS)), (ROWS - 1, COLS - 1)), shape=(350, 399)).toarray()
    x = rng.random(399)
    assert np.allclose(eng.spmm(e1, x)[:, 0], A @ x, rtol=1e-14, atol=0)
    assert np.allclose(eng.ata_mul(e1, x, 0.1), A.T @ (A @ x) + 0.1 * x, rtol=1e-14, atol=0)
    rhs = rng.random((399, D))
    want = np.linalg.solve(A.T @ A + 0.5 * np.eye(399), rhs)
    got, iters = eng.cg_solve(e1, rhs, 0.5)                     # default tol = eps·n, maxiter = n
    assert rel_err(got, want) <= 1e-10
    ref, iters_o = orc.solve_cg2(350, 399, ROWS, COLS, rhs, 0.5, tol=np.finfo(float).eps * 399)
    assert rel_err(got, ref) <= 1e-10
    assert np.all(np.abs(iters - iters_o) <= 2)
    got6, it6 = eng.cg_solve(e1, rhs, 0.5, tol=1e-6)            # test/heavy_copyto.jl:46-50
    ref6, it6o = orc.solve_cg2(350, 399, ROWS, COLS, rhs, 0.5, tol=1e-6)
    assert np.array_equal(it6, it6o)                            # same stopping iteration per column
    assert rel_err(got6, ref6) <= 1e-10
    eng.close()


@pytest.mark.parametrize("D", [5, 32])
def test_sample_beta_lambda_beta_and_uhat_match_oracle(D):
    rng = np.random.default_rng(40 + D)
    N, numF = 500, 120
    dens = rng.random((N, numF)) < 0.08
    r, c = np.nonzero(dens)
    rows, cols = (r + 1).astype(np.int32), (c + 1).astype(np.int32)
    eng, e1, e2, rel, (ids, vals), _ = engine_with_features(D, rows, cols, N, numF, seed=D)
    U = rng.standard_normal((N, D))
    eng.set_factors(e1, U)
    mu = rng.standard_normal(D) * 0.3
    G = rng.standard_normal((D, D)) * 0.2
    Lambda = G @ G.T + 2.0 * np.eye(D)
    E1, E2 = rng.standard_normal((N, D)), rng.standard_normal((numF, D))
    lb = 4.0
    beta, rhs, iters = eng.sample_beta(e1, mu, Lambda, lb, E1=E1, E2=E2, want_rhs=True)
    N1, N2 = orc.color_noise(Lambda, E1), orc.color_noise(Lambda, E2)
    rhs_o = orc.beta_rhs_sbm(U, mu, N1, N2, rows, cols, numF, lb)
    assert rel_err(rhs, rhs_o) <= 1e-12
    beta_o, iters_o = orc.solve_cg2(N, numF, rows, cols, rhs_o, lb, tol=np.finfo(float).eps * numF)
    assert rel_err(beta, beta_o) <= 1e-10, rel_err(beta, beta_o)
    assert np.all(np.abs(iters - iters_o) <= 2)
    assert rel_err(eng.get_beta(e1), beta) == 0.0
    # beta' * beta and lambda_beta with the injected Gamma variate
    BtB = eng.beta_gram(e1)
    assert rel_err(BtB, orc.btb(beta_o)) <= 1e-10
    lam, shape = eng.sample_lambda_beta(e1, Lambda, 1e-3, 1.0, 0.77)
    lam_o, shape_o = orc.lambda_beta(orc.btb(beta_o), Lambda, numF, 1e-3, 1.0, 0.77)
    assert shape == shape_o and abs(lam - lam_o) <= 1e-10 * lam_o
    # uhat = (F·beta)', per-row mean, half-sweep with the mean matrix, statistics of U − uhat
    uhat = eng.update_uhat(e1, mu, want=True)
    uhat_o = orc.f_mul_beta_sbm(N, numF, rows, cols, beta_o)
    assert rel_err(uhat, uhat_o) <= 1e-10
    n, NU, NS = eng.nw_stats_uhat(e1)
    n_o, NU_o, NS_o = orc.nw_stats(U, uhat)
    assert n == n_o and rel_err(NU, NU_o) <= 1e-10 and rel_err(NS, NS_o) <= 1e-10
    V = rng.standard_normal((7, D)) * 0.3
    eng.set_factors(e2, V)
    eng.set_relation_params(rel, 2.0, float(vals.mean()))
    Z = rng.standard_normal((N, D))
    eng.sample_mode_uhat(e1, Lambda, Z)
    idf = orc.FastIDF(ids, vals, [N, 7])
    Uo = [U.copy(), V.copy()]
    orc.sample_latent_all(idf, 0, Uo, 2.0, float(vals.mean()), mu + uhat, Lambda, Z)
    assert rel_err(eng.get_factors(e1), Uo[0]) <= 1e-10
    eng.close()


def test_philox_beta_draw_has_the_right_moments():
    """Without injected noise the rhs noise is device Philox: beta must still solve its own rhs."""
    rng = np.random.default_rng(5)
    N, numF, D = 400, 60, 8
    dens = rng.random((N, numF)) < 0.1
    r, c = np.nonzero(dens)
    rows, cols = (r + 1).astype(np.int32), (c + 1).astype(np.int32)
    eng, e1, *_ = engine_with_features(D, rows, cols, N, numF)
    eng.set_factors(e1, rng.standard_normal((N, D)))
    beta, rhs, iters = eng.sample_beta(e1, np.zeros(D), np.eye(D), 2.0, want_rhs=True)
    F = dens.astype(float)
    assert rel_err((F.T @ F + 2.0 * np.eye(numF)) @ beta, rhs) <= 1e-9
    lam, shape = eng.sample_lambda_beta(e1, np.eye(D), 1e-3, 1.0)
    assert lam > 0 and shape == (1e-3 + numF * D) / 2
    eng.close()


def test_macau_with_and_without_features_end_to_end():
    """The reference's smoke tests (test/basic.jl:104-107, test/parallel_latent_basic.jl:4-17) through the Python mirror of
    macau(): BPMF recovers planted low-rank data; Macau with sparse-binary side features runs the beta / lambda_beta path."""
    import scipy.sparse as sp

    import bdf_b200

    rng = np.random.default_rng(6)
    N1, N2, D0 = 300, 200, 3
    A, B = rng.standard_normal((N1, D0)), rng.standard_normal((N2, D0))
    mask = rng.random((N1, N2)) < 0.25
    W = sp.csc_matrix(np.where(mask, A @ B.T + 0.3 * rng.standard_normal((N1, N2)), 0.0))
    rd = bdf_b200.RelationData(W, class_cut=0.0, alpha=5.0)
    bdf_b200.assignToTest(rd.relations[0], 2000, rng)
    res = bdf_b200.macau(rd, burnin=20, psamples=20, num_latent=6, verbose=False, seed=3)
    assert res["RMSE"] < 0.6 and 0.5 < res["ROC"] <= 1.0
    assert res["predictions"]["pred"].shape == (2000,) and np.all(res["predictions"]["stdev"] >= 0)
    assert rd.entities[0].model.sample.shape == (N1, 6)
    # side features that carry the row factors: 40 random binary columns correlated with A
    feat = (A @ rng.standard_normal((D0, 40)) + 0.5 * rng.standard_normal((N1, 40)) > 0.8)
    r, c = np.nonzero(feat)
    F = bdf_b200.SparseBinMatrix(r + 1, c + 1, N1, 40)
    rd2 = bdf_b200.RelationData(W, class_cut=0.0, alpha=5.0, feat1=F)
    bdf_b200.assignToTest(rd2.relations[0], 2000, rng)
    res2 = bdf_b200.macau(rd2, burnin=15, psamples=15, num_latent=6, verbose=False, seed=4, host_noise=np.random.default_rng(9))
    assert res2["RMSE"] < 0.7
    assert rd2.entities[0].model.beta.shape == (40, 6) and np.all(np.isfinite(rd2.entities[0].model.beta))
    assert rd2.entities[0].lambda_beta > 0


def test_long_rows_are_chunked_deterministically():
    """A popular feature bit (one column with 5000 entries > the 1024-index chunk) is summed as chunk partials added in
    chunk order: deterministic, and equal to the reference's sequential sum to rounding."""
    rng = np.random.default_rng(77)
    m, n, D = 6000, 50, 32
    rows = np.concatenate([rng.permutation(m)[:5000] + 1, rng.integers(1, m + 1, 3000)]).astype(np.int32)
    cols = np.concatenate([np.full(5000, 7), rng.integers(1, n + 1, 3000)]).astype(np.int32)
    eng, e1, *_ = engine_with_features(D, rows, cols, m, n, seed=3)
    Xt = rng.standard_normal((m, D))
    Y1 = eng.spmm(e1, Xt, transpose=True)
    Y2 = eng.spmm(e1, Xt, transpose=True)
    assert np.array_equal(Y1, Y2)
    for d in (0, 5, 31):
        want = orc.sbm_tmul(m, n, rows, cols, Xt[:, d])
        assert rel_err(Y1[:, d], want) <= 1e-13
        short = np.ones(n, dtype=bool)
        short[6] = False
        assert np.array_equal(Y1[short, d], want[short])   # rows within one chunk stay bit-exact
    eng.close()


@pytest.mark.parametrize("D", [5, 32])
def test_general_sparse_features_match_julia_summation_order(D):
    """Entity(F = sprand(...)) — the SparseMatrixCSC features of test/parallel_latent_basic.jl:4 and test/parallel_mult.jl:4-18.
    scipy's CSC/CSR matvecs accumulate in the same order as Julia's (ascending column per row for F*x, stored order per
    column for F'x) with un-fused multiply-add, so the device products must match them bit for bit."""
    import scipy.sparse as sp

    import bdf_b200

    rng = np.random.default_rng(300 + D)
    N, numF = 50, 20
    X = sp.random(N, numF, 0.1, random_state=int(D), format="csc")   # sprand(50, 20, 0.1)
    X.sort_indices()
    eng = bdf_b200.Engine(D)
    e1, e2 = eng.add_entity(N), eng.add_entity(10)
    ids = np.stack([rng.integers(1, N + 1, 300), rng.integers(1, 11, 300)], axis=1)
    vals = rng.standard_normal(300)
    rel = eng.add_relation([e1, e2], ids, vals)
    eng.set_features(e1, X)
    B = rng.standard_normal((numF, D))
    T = rng.standard_normal((N, D))
    Y = eng.spmm(e1, B)
    Yt = eng.spmm(e1, T, transpose=True)
    Xcsr_t = sp.csr_matrix(X.T)
    for d in range(D):
        assert np.array_equal(Y[:, d], X @ B[:, d])            # Z2 = X * Y2, test/parallel_mult.jl:8
        assert np.array_equal(Yt[:, d], Xcsr_t @ T[:, d])      # Z1 = X' * Y1, test/parallel_mult.jl:7
    x = rng.standard_normal(numF)
    Xd = X.toarray()
    assert rel_err(eng.ata_mul(e1, x, 0.5), Xd.T @ (Xd @ x) + 0.5 * x) <= 1e-14
    rhs = rng.standard_normal((numF, D))
    got, iters = eng.cg_solve(e1, rhs, 0.75, tol=1e-12, maxiter=500)
    assert rel_err(got, np.linalg.solve(Xd.T @ Xd + 0.75 * np.eye(numF), rhs)) <= 1e-9
    # the reference's smoke test itself (test/parallel_latent_basic.jl:4-15): rank-2 data, sparse real features, D=5
    A2, B2 = rng.standard_normal((N, 2)), rng.standard_normal((10, 2))
    rd = bdf_b200.RelationData(sp.csc_matrix(A2 @ B2.T), class_cut=0.5, feat1=X)
    bdf_b200.assignToTest(rd.relations[0], 50, rng)
    res = bdf_b200.macau(rd, burnin=3, psamples=3, num_latent=5, verbose=False)
    assert rd.entities[0].model.beta.shape == (numF, 5) and np.isfinite(res["RMSE"])
    eng.close()


@pytest.mark.parametrize("partition", ["cyclic", "balanced"])
def test_feature_path_on_sharded_entities(partition):
    """world = 3 handles on one GPU: every rank keeps the whole F and beta, draws the same beta from its replica of U, files uhat by
    slot, samples only its rows and reduces only its rows' statistics (the caller all-reduces them)."""
    import bdf_b200
    from bdf_b200.shard import balanced_partition

    rng = np.random.default_rng(77)
    N, M, numF, D, W = 90, 11, 40, 16, 3
    dens = rng.random((N, numF)) < 0.15
    r, c = np.nonzero(dens)
    rows, cols = (r + 1).astype(np.int32), (c + 1).astype(np.int32)
    nnz = 700
    ids = np.stack([rng.integers(1, N + 1, nnz), rng.integers(1, M + 1, nnz)], 1).astype(np.int64)
    ids[:200, 0] = 9
    vals = rng.standard_normal(nnz)
    mean = float(vals.mean())
    U, V = rng.standard_normal((N, D)) * 0.5, rng.standard_normal((M, D)) * 0.5
    beta0 = rng.standard_normal((numF, D)) * 0.2
    mu = rng.standard_normal(D) * 0.3
    G = rng.standard_normal((D, D)) * 0.2
    Lambda = G @ G.T + 2.0 * np.eye(D)
    maps = [balanced_partition(np.bincount(ids[:, 0] - 1, minlength=N), W, 10.0), balanced_partition(np.bincount(ids[:, 1] - 1, minlength=M), W, 10.0)]
    owner = maps[0] if partition == "balanced" else np.arange(N) % W
    E1, E2, Z = rng.standard_normal((N, D)), rng.standard_normal((numF, D)), rng.standard_normal((N, D))
    uhat_o = orc.f_mul_beta_sbm(N, numF, rows, cols, beta0)
    Uo = [U.copy(), V.copy()]
    orc.sample_latent_all(orc.FastIDF(ids, vals, [N, M]), 0, Uo, 2.0, mean, mu + uhat_o, Lambda, Z)
    rhs_o = orc.beta_rhs_sbm(U, mu, orc.color_noise(Lambda, E1), orc.color_noise(Lambda, E2), rows, cols, numF, 3.0)
    beta_o, _ = orc.solve_cg2(N, numF, rows, cols, rhs_o, 3.0, tol=np.finfo(float).eps * numF)
    got_rows = U.copy()
    tot = [0.0, np.zeros(D), np.zeros((D, D))]
    for rk in range(W):
        eng = bdf_b200.Engine(D, rank=rk, world=W)
        if partition == "balanced":
            e1, e2 = eng.add_entity_partitioned(N, maps[0]), eng.add_entity_partitioned(M, maps[1])
        else:
            e1, e2 = eng.add_entity(N), eng.add_entity(M)
        rel = eng.add_relation([e1, e2], ids, vals)
        eng.set_relation_params(rel, 2.0, mean)
        eng.set_features(e1, bdf_b200.SparseBinMatrix(rows, cols, N, numF))
        eng.set_factors(e1, U)
        eng.set_factors(e2, V)
        eng.set_beta(e1, beta0)
        assert rel_err(eng.update_uhat(e1, mu, want=True), uhat_o) <= 1e-12
        n, NU, NS = eng.nw_stats_uhat(e1)               # this rank's rows only
        tot[0] += n; tot[1] += NU; tot[2] += NS
        beta, rhs, _ = eng.sample_beta(e1, mu, Lambda, 3.0, E1=E1, E2=E2, want_rhs=True)   # from the full replica: same on every rank
        assert rel_err(rhs, rhs_o) <= 1e-12 and rel_err(beta, beta_o) <= 1e-10
        eng.set_beta(e1, beta0)
        eng.update_uhat(e1, mu)
        eng.sample_mode_uhat(e1, Lambda, Z)
        mine = owner == rk
        got_rows[mine] = eng.get_factors(e1)[mine]
        eng.close()
    n_o, NU_o, NS_o = orc.nw_stats(U, uhat_o)
    assert tot[0] == n_o and rel_err(tot[1], NU_o) <= 1e-10 and rel_err(tot[2], NS_o) <= 1e-10
    assert rel_err(got_rows, Uo[0]) <= 1e-10
