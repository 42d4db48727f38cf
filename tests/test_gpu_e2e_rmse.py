"""End-to-end check (BASELINE.json north_star, part 2 of correctness): the posterior-mean test RMSE of the CUDA engine
matches the restated reference's (the CPU oracle driven through the same Gibbs loop, src/macau.jl:80-203) within a
stated tolerance. The two runs use different random streams (device Philox vs numpy), so the tolerance is statistical:
|ΔRMSE| ≤ 0.03 on this planted problem (noise floor 1/sqrt(alpha) ≈ 0.447, mean predictor ≈ 1.4)."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TOL_RMSE = 0.03


def oracle_macau(ids, vals, test_ids, test_vals, dims, D, alpha, burnin, psamples, seed):
    rng = np.random.default_rng(seed)
    idf = orc.FastIDF(ids, vals, dims)
    mean = float(vals.mean())
    U = [np.zeros((d, D)) for d in dims]
    mu = [np.zeros(D) for _ in dims]
    Lam = [5.0 * np.eye(D) for _ in dims]          # reset!, src/RelationData.jl:66-90
    acc = np.zeros(len(test_vals))
    for it in range(burnin + psamples):
        for m in range(len(dims)):                 # src/macau.jl:96-134
            orc.sample_latent_all(idf, m, U, alpha, mean, mu[m], Lam[m], rng.standard_normal((dims[m], D)), nshards=4)
            N, NU, NS = orc.nw_stats(U[m])
            mu_N, beta_N, T_N, nu_N = orc.cond_normal_wishart(N, NU, NS, np.zeros(D), 2.0, np.eye(D), float(D))
            mu[m], Lam[m] = orc.nw_rand(mu_N, beta_N, T_N, orc.bartlett_factor(rng, D, nu_N), rng.standard_normal(D))
        if it >= burnin:
            acc += orc.pred(test_ids, U, mean)     # src/macau.jl:143-175
    return float(np.sqrt(np.mean((acc / psamples - test_vals) ** 2)))


@pytest.mark.parametrize("D", [8, 32])
def test_posterior_mean_rmse_matches_the_restated_reference(D):
    import bdf_b200

    rng = np.random.default_rng(2024)
    N1, N2, D0, alpha = 500, 300, 4, 5.0
    A, B = rng.standard_normal((N1, D0)), rng.standard_normal((N2, D0)) * 0.7
    nnz = 30000
    ids = np.stack([rng.integers(1, N1 + 1, nnz), rng.integers(1, N2 + 1, nnz)], axis=1)
    vals = np.einsum("ij,ij->i", A[ids[:, 0] - 1], B[ids[:, 1] - 1]) + rng.standard_normal(nnz) / np.sqrt(alpha)
    ntest = 3000
    te, tr = slice(0, ntest), slice(ntest, None)
    burnin, psamples = 40, 40
    rmse_cpu = oracle_macau(ids[tr], vals[tr], ids[te], vals[te], [N1, N2], D, alpha, burnin, psamples, seed=1)
    rd = bdf_b200.RelationData((ids[tr], vals[tr], [N1, N2]), alpha=alpha, class_cut=0.0)
    bdf_b200.setTest(rd.relations[0], ids[te], vals[te])
    res = bdf_b200.macau(rd, num_latent=D, burnin=burnin, psamples=psamples, verbose=False, seed=7)
    assert abs(res["RMSE"] - rmse_cpu) <= TOL_RMSE, (res["RMSE"], rmse_cpu)
    assert res["RMSE"] < 0.6
    # host-injected Normal-Wishart variates (the reference keeps that stream on the host) give the same quality
    rd2 = bdf_b200.RelationData((ids[tr], vals[tr], [N1, N2]), alpha=alpha, class_cut=0.0)
    bdf_b200.setTest(rd2.relations[0], ids[te], vals[te])
    res2 = bdf_b200.macau(rd2, num_latent=D, burnin=burnin, psamples=psamples, verbose=False, seed=8, host_noise=np.random.default_rng(3))
    assert abs(res2["RMSE"] - rmse_cpu) <= TOL_RMSE, (res2["RMSE"], rmse_cpu)
