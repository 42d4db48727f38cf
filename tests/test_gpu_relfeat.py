"""GPU parity of relation-level features (Relation.F — one feature row per observation): sample_beta_rel (src/sampling.jl:322-337), the
per-observation offset linear_values in the row draws (src/sampling.jl:273) and in pred (src/sampling.jl:9-19), against the oracle."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("K,D", [(2, 10), (3, 30)])
def test_sample_beta_rel_offsets_and_prediction_match_oracle(K, D):
    import bdf_b200

    rng = np.random.default_rng(100 * K + D)
    dims = [40, 23, 6][:K]
    nnz, nF = 1500, 7
    ids = np.stack([rng.integers(1, d + 1, nnz) for d in dims], 1).astype(np.int64)
    ids[:700, 0] = 3
    vals = rng.standard_normal(nnz) + 0.3
    F = rng.standard_normal((nnz, nF))
    U = [rng.standard_normal((d, D)) * 0.4 for d in dims]
    alpha, lam, mean = 2.5, 1.7, float(vals.mean())
    eng = bdf_b200.Engine(D)
    ents = [eng.add_entity(d) for d in dims]
    rel = eng.add_relation(ents, ids, vals)
    eng.set_relation_params(rel, alpha, mean)
    eng.set_relation_features(rel, F)
    for e, u in zip(ents, U):
        eng.set_factors(e, u)
    # beta = 0 right after registration: linear_values = mean_value, so the draws equal the feature-less ones
    Z = rng.standard_normal((dims[0], D))
    G = rng.standard_normal((D, D)) * 0.3
    Lambda, mu = G @ G.T + 2.0 * np.eye(D), rng.standard_normal(D) * 0.1
    eng.sample_mode(ents[0], mu, Lambda, Z)
    Uo = [u.copy() for u in U]
    orc.sample_latent_all(orc.FastIDF(ids, vals, dims), 0, Uo, alpha, mean, mu, Lambda, Z)
    assert rel_err(eng.get_factors(ents[0]), Uo[0]) <= 1e-10
    eng.set_factors(ents[0], U[0])
    # sample_beta_rel with injected normals
    z1, z2 = rng.standard_normal(nnz), rng.standard_normal(nF)
    beta = eng.sample_beta_rel(rel, lam, z1, z2)
    udot = orc.pred(ids, U, 0.0)
    beta_o = orc.sample_beta_rel(F, vals, udot, mean, alpha, lam, z1, z2)
    assert rel_err(beta, beta_o) <= 1e-10
    lin = mean + F @ beta_o   # r.temp.linear_values
    # training residuals for sample_alpha use pred(r) = udot + linear_values
    sse, n = eng.train_sse(rel)
    err = udot + lin - vals
    assert n == nnz and abs(sse - float(err @ err)) <= 1e-11 * float(err @ err)
    # row draws of every mode with the per-observation offsets
    for mode in range(K):
        Zm = rng.standard_normal((dims[mode], D))
        eng.sample_mode(ents[mode], mu, Lambda, Zm)
        got = eng.get_factors(ents[mode])
        eng.set_factors(ents[mode], U[mode])
        others = [m for m in range(K) if m != mode]
        worst = 0.0
        for i in range(dims[mode]):
            sel = np.nonzero(ids[:, mode] == i + 1)[0]
            want = orc.sample_row(D, [{"U": [U[m] for m in others], "ids": [ids[sel, m] for m in others], "vals": vals[sel], "offsets": lin[sel],
                                       "alpha": alpha}], mu, Lambda, Zm[i])
            worst = max(worst, rel_err(got[i], want))
        assert worst <= 1e-10, (mode, worst)
    # pred(r, probe_vec, F) = udot + F*beta + mean
    tids = np.stack([rng.integers(1, d + 1, 60) for d in dims], 1)
    tF = rng.standard_normal((60, nF))
    assert rel_err(eng.predict(rel, tids, tF), orc.pred(tids, U, mean) + tF @ beta_o) <= 1e-12
    with pytest.raises(bdf_b200.BDFError):
        eng.predict(rel, tids)   # a relation with features needs the test features
    # Philox draw solves its own system: K·beta = aFt_y can only be checked statistically; check determinism and finiteness
    b1 = eng.sample_beta_rel(rel, lam)
    b2 = eng.sample_beta_rel(rel, lam)
    assert np.array_equal(b1, b2) and np.all(np.isfinite(b1))
    eng.close()


def test_macau_with_relation_features_recovers_their_effect():
    import bdf_b200
    from bdf_b200.relation_data import Entity, IndexedDF, Relation, assignToTest

    rng = np.random.default_rng(12)
    N, M, D0, nnz, nF = 80, 50, 2, 3000, 3
    A, B = rng.standard_normal((N, D0)), rng.standard_normal((M, D0))
    ids = np.stack([rng.integers(1, N + 1, nnz), rng.integers(1, M + 1, nnz)], 1).astype(np.int64)
    F = rng.standard_normal((nnz, nF))
    b_true = np.array([1.5, -2.0, 0.7])
    vals = np.einsum("ij,ij->i", A[ids[:, 0] - 1], B[ids[:, 1] - 1]) + F @ b_true + 0.2 * rng.standard_normal(nnz)
    out = {}
    for name, feats in (("with", F), ("without", None)):
        a, b = Entity("a"), Entity("b")
        r = Relation(IndexedDF(ids.copy(), vals.copy(), [N, M]), "ab", [a, b], class_cut=0.0, alpha=5.0)
        r.F = feats
        assignToTest(r, 300, np.random.default_rng(1))
        rd = bdf_b200.RelationData()
        rd.addRelation(r)
        res = bdf_b200.macau(rd, num_latent=4, burnin=40, psamples=40, verbose=False, seed=5)
        out[name] = res["RMSE"]
        if feats is not None:
            assert r.F.shape[0] == r.numData() and r.test_F.shape == (300, nF)
            assert np.max(np.abs(r.model.beta - b_true)) < 0.2
    assert out["with"] < 0.5 * out["without"]
