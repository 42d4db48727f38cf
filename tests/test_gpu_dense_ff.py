"""GPU parity of the FF direct-solve path (A9: en.FF = full(FᵀF), solve_full — src/RelationData.jl:337-339,
src/sampling.jl:303-304, 314-320) and of dense feature matrices, against the CPU oracle through the C ABI."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def setup(D, N, numF, F, seed=0, n2=7):
    import bdf_b200

    rng = np.random.default_rng(seed)
    eng = bdf_b200.Engine(D)
    e1, e2 = eng.add_entity(N), eng.add_entity(n2)
    nnz = 5 * N
    ids = np.stack([rng.integers(1, N + 1, nnz), rng.integers(1, n2 + 1, nnz)], axis=1)
    vals = rng.standard_normal(nnz)
    rel = eng.add_relation([e1, e2], ids, vals)
    eng.set_features(e1, F)
    return eng, e1, e2, rel, rng


@pytest.mark.parametrize("D", [5, 32])
def test_ff_of_a_sparse_binary_matrix_is_exact_and_solve_full_matches_oracle(D):
    import bdf_b200

    rng = np.random.default_rng(60 + D)
    N, numF = 400, 75   # numF is not a multiple of D: the last identity block is partial
    dens = rng.random((N, numF)) < 0.1
    r, c = np.nonzero(dens)
    rows, cols = (r + 1).astype(np.int32), (c + 1).astype(np.int32)
    eng, e1, e2, rel, _ = setup(D, N, numF, bdf_b200.SparseBinMatrix(rows, cols, N, numF))
    FF = eng.compute_ff(e1, want=True)
    Fd = dens.astype(np.float64)
    assert np.array_equal(FF, Fd.T @ Fd)  # integer counts: exact
    rhs = rng.standard_normal((numF, D))
    x = eng.solve_full(e1, rhs, 0.5)
    assert rel_err(x, orc.solve_full(FF, rhs, 0.5)) <= 1e-10
    # sample_beta with use_ff: same rhs as the CG path, direct solve for beta
    U = rng.standard_normal((N, D))
    eng.set_factors(e1, U)
    mu = rng.standard_normal(D) * 0.3
    G = rng.standard_normal((D, D)) * 0.2
    Lambda = G @ G.T + 2.0 * np.eye(D)
    E1, E2 = rng.standard_normal((N, D)), rng.standard_normal((numF, D))
    lb = 4.0
    beta, rhs_d, iters = eng.sample_beta(e1, mu, Lambda, lb, E1=E1, E2=E2, want_rhs=True)
    rhs_o = orc.beta_rhs_sbm(U, mu, orc.color_noise(Lambda, E1), orc.color_noise(Lambda, E2), rows, cols, numF, lb)
    assert rel_err(rhs_d, rhs_o) <= 1e-12
    assert rel_err(beta, orc.solve_full(FF, rhs_o, lb)) <= 1e-10
    assert np.all(iters == 0)
    # switching FF off goes back to CG and agrees with the direct solve to the CG tolerance
    eng.set_use_ff(e1, False)
    beta_cg, _, iters = eng.sample_beta(e1, mu, Lambda, lb, E1=E1, E2=E2, want_rhs=True)
    assert np.all(iters > 0) and rel_err(beta_cg, beta) <= 1e-9
    eng.close()


@pytest.mark.parametrize("D", [8, 30])
def test_dense_features_products_ff_and_beta(D):
    rng = np.random.default_rng(70 + D)
    N, numF = 300, 29   # MovieLens-like: a few dozen dense feature columns
    F = rng.standard_normal((N, numF))
    eng, e1, e2, rel, _ = setup(D, N, numF, F)
    X = rng.standard_normal((numF, D))
    assert rel_err(eng.spmm(e1, X), F @ X) <= 1e-13
    T = rng.standard_normal((N, D))
    assert rel_err(eng.spmm(e1, T, transpose=True), F.T @ T) <= 1e-13
    x = rng.standard_normal(numF)
    assert rel_err(eng.ata_mul(e1, x, 0.5), orc.dense_ata_mul(F, x, 0.5)) <= 1e-12   # test/solver.jl:14-20
    FF = eng.compute_ff(e1, want=True)
    assert rel_err(FF, F.T @ F) <= 1e-13
    rhs = rng.standard_normal((numF, D))
    assert rel_err(eng.solve_full(e1, rhs, 0.5), orc.solve_full(F.T @ F, rhs, 0.5)) <= 1e-10   # test/solver.jl:4-11
    # CG on the dense operator against the direct solve (test/parallel_matrix.jl:107-109 pattern)
    eng.set_use_ff(e1, False)
    xcg, iters = eng.cg_solve(e1, rhs, 0.5)
    assert rel_err(xcg, orc.solve_full(F.T @ F, rhs, 0.5)) <= 1e-9
    eng.set_use_ff(e1, True)
    # beta draw + uhat
    U = rng.standard_normal((N, D))
    eng.set_factors(e1, U)
    mu, Lambda = rng.standard_normal(D) * 0.3, 2.0 * np.eye(D)
    E1, E2 = rng.standard_normal((N, D)), rng.standard_normal((numF, D))
    beta, rhs_d, _ = eng.sample_beta(e1, mu, Lambda, 3.0, E1=E1, E2=E2, want_rhs=True)
    N1, N2 = orc.color_noise(Lambda, E1), orc.color_noise(Lambda, E2)
    rhs_o = F.T @ ((U - mu) + N1) + np.sqrt(3.0) * N2   # src/sampling.jl:300
    assert rel_err(rhs_d, rhs_o) <= 1e-12
    assert rel_err(beta, orc.solve_full(F.T @ F, rhs_o, 3.0)) <= 1e-10
    assert rel_err(eng.update_uhat(e1, mu, want=True), F @ beta) <= 1e-12
    eng.close()


def test_macau_takes_the_ff_path_for_small_feature_matrices():
    """compute_ff_size selects solve_full vs CG exactly as reset! does (src/RelationData.jl:336-343); both recover the signal."""
    import bdf_b200

    rng = np.random.default_rng(11)
    N, M, numF, D0 = 300, 80, 12, 2
    F = rng.standard_normal((N, numF))
    B0 = rng.standard_normal((numF, D0))
    Uo, Vo = F @ B0 * 0.5, rng.standard_normal((M, D0))
    mask = rng.random((N, M)) < 0.15
    Y = sp.csc_matrix(np.where(mask, Uo @ Vo.T + 0.1 * rng.standard_normal((N, M)), 0.0))
    out = {}
    for name, ff_size in (("ff", 6500), ("cg", 0)):
        rd = bdf_b200.RelationData(Y, feat1=F, class_cut=0.0, alpha=5.0)
        bdf_b200.assignToTest(rd.relations[0], 300, np.random.default_rng(3))
        res = bdf_b200.macau(rd, num_latent=4, burnin=40, psamples=40, verbose=False, compute_ff_size=ff_size, seed=1)
        assert rd.entities[0].use_FF == (name == "ff")
        out[name] = res["RMSE"]
    base = float(np.std(Y.data))
    assert out["ff"] < 0.5 * base and out["cg"] < 0.5 * base
    assert abs(out["ff"] - out["cg"]) < 0.15 * base
