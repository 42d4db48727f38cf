"""GPU parity for entities that sit in several relations (sample_user2_all! / sample_user2, src/sampling.jl:251-289):
the row's precision and rhs are sums over the relations, each with its own alpha and mean. Oracle: orc.sample_row."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-10


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def oracle_rows(D, n_rows, uses, mu, Lambda, Z):
    """uses: list of (ids, vals, mode, [partner factor matrices in mode order], alpha, mean)."""
    out = np.zeros((n_rows, D))
    for i in range(n_rows):
        rels = []
        for ids, vals, mode, Us, alpha, mean in uses:
            sel = np.nonzero(ids[:, mode] == i + 1)[0]  # table order, as FastIDF keeps it
            others = [m for m in range(ids.shape[1]) if m != mode]
            rels.append({"U": Us, "ids": [ids[sel, m] for m in others], "vals": vals[sel], "offset": mean, "alpha": alpha})
        out[i] = orc.sample_row(D, rels, mu, Lambda, Z[i])
    return out


@pytest.mark.parametrize("D", [5, 32, 100])
def test_entity_in_two_matrix_relations(D):
    import bdf_b200

    rng = np.random.default_rng(900 + D)
    nA, nB, nC = 40, 23, 17
    idsAB = np.stack([rng.integers(1, nA + 1, 600), rng.integers(1, nB + 1, 600)], 1).astype(np.int64)
    idsAC = np.stack([rng.integers(1, nA + 1, 300), rng.integers(1, nC + 1, 300)], 1).astype(np.int64)
    idsAB[idsAB[:, 0] == 3, 0] = 4   # row 3 of A has no AB observations
    idsAC[idsAC[:, 0] == 3, 0] = 5   # ... and none in AC either: prior draw
    idsAC[idsAC[:, 0] == 7, 0] = 8   # row 7: only AB
    vAB, vAC = rng.standard_normal(600), rng.standard_normal(300) + 1.0
    UA, UB, UC = (rng.standard_normal((n, D)) * 0.3 for n in (nA, nB, nC))
    G = rng.standard_normal((D, D)) * 0.3
    Lambda, mu = G @ G.T + 2.0 * np.eye(D), rng.standard_normal(D) * 0.1
    eng = bdf_b200.Engine(D)
    eA, eB, eC = eng.add_entity(nA), eng.add_entity(nB), eng.add_entity(nC)
    rAB = eng.add_relation([eA, eB], idsAB, vAB)
    rAC = eng.add_relation([eA, eC], idsAC, vAC)
    eng.set_relation_params(rAB, 2.0, float(vAB.mean()))
    eng.set_relation_params(rAC, 0.7, float(vAC.mean()))
    for e, u in ((eA, UA), (eB, UB), (eC, UC)):
        eng.set_factors(e, u)
    Z = rng.standard_normal((nA, D))
    eng.sample_mode(eA, mu, Lambda, Z)
    got = eng.get_factors(eA)
    want = oracle_rows(D, nA, [(idsAB, vAB, 0, [UB], 2.0, float(vAB.mean())), (idsAC, vAC, 0, [UC], 0.7, float(vAC.mean()))], mu, Lambda, Z)
    assert rel_err(got, want) <= TOL
    # B sits in one relation only and must be untouched by the merged path
    ZB = rng.standard_normal((nB, D))
    eng.set_factors(eA, UA)
    eng.sample_mode(eB, mu, Lambda, ZB)
    wantB = oracle_rows(D, nB, [(idsAB, vAB, 1, [UA], 2.0, float(vAB.mean()))], mu, Lambda, ZB)
    assert rel_err(eng.get_factors(eB), wantB) <= TOL
    eng.close()


@pytest.mark.parametrize("D", [10, 30])
def test_entity_in_a_matrix_and_a_tensor_relation_with_a_long_row(D):
    import bdf_b200

    rng = np.random.default_rng(950 + D)
    nA, nB, nC, nE = 12, 300, 9, 11
    nAB = 30000
    idsAB = np.stack([rng.integers(1, nA + 1, nAB), rng.integers(1, nB + 1, nAB)], 1).astype(np.int64)
    idsAB[:20000, 0] = 2                                    # a row long enough to be chunked inside its relation
    idsT = np.stack([rng.integers(1, nA + 1, 900), rng.integers(1, nC + 1, 900), rng.integers(1, nE + 1, 900)], 1).astype(np.int64)
    vAB, vT = rng.standard_normal(nAB), rng.standard_normal(900)
    UA, UB, UC, UE = (rng.standard_normal((n, D)) * 0.3 for n in (nA, nB, nC, nE))
    G = rng.standard_normal((D, D)) * 0.3
    Lambda, mu = G @ G.T + 2.0 * np.eye(D), rng.standard_normal(D) * 0.1
    eng = bdf_b200.Engine(D)
    eA, eB, eC, eE = (eng.add_entity(n) for n in (nA, nB, nC, nE))
    rAB = eng.add_relation([eA, eB], idsAB, vAB)
    rT = eng.add_relation([eC, eA, eE], idsT[:, [1, 0, 2]], vT)   # A is the middle mode of the tensor
    eng.set_relation_params(rAB, 1.5, 0.1)
    eng.set_relation_params(rT, 3.0, -0.2)
    for e, u in ((eA, UA), (eB, UB), (eC, UC), (eE, UE)):
        eng.set_factors(e, u)
    Z = rng.standard_normal((nA, D))
    eng.sample_mode(eA, mu, Lambda, Z)
    got = eng.get_factors(eA)
    want = oracle_rows(D, nA, [(idsAB, vAB, 0, [UB], 1.5, 0.1), (idsT[:, [1, 0, 2]], vT, 1, [UC, UE], 3.0, -0.2)], mu, Lambda, Z)
    assert rel_err(got, want) <= TOL
    # deterministic: the split partials are added in item order
    eng.set_factors(eA, UA)
    eng.sample_mode(eA, mu, Lambda, Z)
    assert np.array_equal(got, eng.get_factors(eA))
    eng.close()


def test_train_sse_and_alpha_draw_match_oracle():
    """sample_alpha (src/macau.jl:84-88, src/sampling.jl:129-134) with the chi-square variate injected."""
    import bdf_b200

    rng = np.random.default_rng(77)
    for dims, D in (([50, 30], 10), ([20, 9, 6], 30)):
        nnz = 3000
        ids = np.stack([rng.integers(1, d + 1, nnz) for d in dims], 1).astype(np.int64)
        ids[:2500, 0] = 7  # one long row
        vals = rng.standard_normal(nnz) + 0.5
        U = [rng.standard_normal((d, D)) * 0.4 for d in dims]
        eng = bdf_b200.Engine(D)
        ents = [eng.add_entity(d) for d in dims]
        rel = eng.add_relation(ents, ids, vals)
        mean = float(vals.mean())
        eng.set_relation_params(rel, 2.0, mean)
        for e, u in zip(ents, U):
            eng.set_factors(e, u)
        err = orc.pred(ids, U, mean) - vals
        sse, n = eng.train_sse(rel)
        assert n == nnz
        assert abs(sse - float(err @ err)) <= 1e-11 * float(err @ err)
        chi2 = float(rng.chisquare(2.0 + nnz))
        got = eng.sample_alpha(rel, 1.0, 2.0, sse, n, chi2)
        want = orc.sample_alpha(1.0, 2.0, err, chi2)
        assert abs(got - want) <= 1e-11 * want
        # Philox draw: alpha concentrates around n / sse
        draws = []
        for _ in range(50):
            draws.append(eng.sample_alpha(rel, 1.0, 2.0, sse, n))
            eng.advance_sweep()
        assert abs(np.mean(draws) / ((2.0 + nnz) / (1.0 + sse)) - 1.0) < 0.02 and np.std(draws) > 0
        eng.close()


def test_macau_with_two_relations_and_alpha_sampling():
    """Driver level: an entity shared by two relations (docs/index.md:200-233 pattern) with alpha sampled."""
    import bdf_b200
    from bdf_b200.relation_data import Entity, IndexedDF, Relation, assignToTest

    rng = np.random.default_rng(5)
    nA, nB, nC, D0 = 120, 60, 40, 3
    A, B, Cm = (rng.standard_normal((n, D0)) for n in (nA, nB, nC))

    def table(X, Y, nnz, noise):
        i = np.stack([rng.integers(1, X.shape[0] + 1, nnz), rng.integers(1, Y.shape[0] + 1, nnz)], 1).astype(np.int64)
        v = np.einsum("ij,ij->i", X[i[:, 0] - 1], Y[i[:, 1] - 1]) + noise * rng.standard_normal(nnz)
        return IndexedDF(i, v, [X.shape[0], Y.shape[0]])

    a, b, c = Entity("a"), Entity("b"), Entity("c")
    r1 = Relation(table(A, B, 4000, 0.3), "ab", [a, b], class_cut=0.0, alpha=1.0)
    r2 = Relation(table(A, Cm, 3000, 0.3), "ac", [a, c], class_cut=0.0, alpha=1.0)
    r1.model.alpha_sample = True
    r2.model.alpha_sample = True
    assignToTest(r1, 400, rng)
    rd = bdf_b200.RelationData()
    rd.addRelation(r1)
    rd.addRelation(r2)
    res = bdf_b200.macau(rd, num_latent=6, burnin=30, psamples=30, verbose=False, seed=3)
    base = float(np.sqrt(np.mean((r1.test_values - r1.data.values.mean()) ** 2)))
    assert np.isfinite(res["RMSE"]) and res["RMSE"] < 0.45 * base
    # the sampled precisions settle near the planted noise level 1/0.3^2 ≈ 11
    assert 6.0 < r1.model.alpha < 18.0 and 6.0 < r2.model.alpha < 18.0


def test_pred_all_full_prediction_and_output_dumps(tmp_path):
    """pred_all (src/sampling.jl:72-97, test/basic.jl:168-174 identity: pred_all[i, j] == pred of that cell), macau(full_prediction,
    output, output_beta) — src/macau.jl:145-162, 228-230."""
    import scipy.sparse as sp

    import bdf_b200
    from bdf_b200 import data_reading as dr

    rng = np.random.default_rng(8)
    N, M, D = 40, 25, 6
    U, V = rng.standard_normal((N, D)), rng.standard_normal((M, D))
    eng = bdf_b200.Engine(D)
    e1, e2 = eng.add_entity(N), eng.add_entity(M)
    ids = np.stack([rng.integers(1, N + 1, 200), rng.integers(1, M + 1, 200)], 1)
    rel = eng.add_relation([e1, e2], ids, rng.standard_normal(200))
    eng.set_relation_params(rel, 2.0, 0.25)
    eng.set_factors(e1, U)
    eng.set_factors(e2, V)
    full = eng.predict_all(rel, (N, M))
    assert rel_err(full, U @ V.T + 0.25) <= 1e-13
    cells = np.stack([rng.integers(1, N + 1, 30), rng.integers(1, M + 1, 30)], 1)
    assert rel_err(full[cells[:, 0] - 1, cells[:, 1] - 1], eng.predict(rel, cells)) <= 1e-13
    eng.close()
    # driver level
    F = sp.random(N, 9, 0.4, random_state=4, format="csc")
    F.data[:] = 1.0
    Y = sp.csc_matrix(np.where(rng.random((N, M)) < 0.4, U[:, :2] @ V[:, :2].T, 0.0))
    rd = bdf_b200.RelationData(Y, feat1=F, class_cut=0.0, alpha=5.0)
    out = str(tmp_path / "run")
    res = bdf_b200.macau(rd, num_latent=4, burnin=3, psamples=4, verbose=False, full_prediction=True, output=out, output_beta=True, output_type="binary", seed=2)
    assert res["predictions_full"].shape == (N, M) and np.all(np.isfinite(res["predictions_full"]))
    S = dr.read_binary_float32(f"{out}-E1-4.binary")           # the last sample, as Float32, num_latent × count
    assert S.shape == (4, N) and np.allclose(S.T, rd.entities[0].model.sample, rtol=1e-6, atol=1e-6)
    B = dr.read_binary_float32(f"{out}-E1-4.beta.binary")
    assert B.shape == (9, 4) and np.allclose(B, rd.entities[0].model.beta, rtol=1e-6, atol=1e-6)


def test_pred_all_with_an_explicit_shard_map_and_emulated_ranks():
    """pred_all (src/sampling.jl:72-97) reads the factor rows back in the caller's order when the entity was created with an explicit shard map
    or the handle is one rank of several (every rank holds all rows, in slot order)."""
    import bdf_b200

    rng = np.random.default_rng(6)
    N, M, D = 37, 23, 12
    U, V = rng.standard_normal((N, D)), rng.standard_normal((M, D))
    ids = np.stack([rng.integers(1, N + 1, 300), rng.integers(1, M + 1, 300)], 1)
    vals = rng.standard_normal(300)
    want = U @ V.T + 0.4
    for rank, world in ((0, 1), (1, 3)):
        eng = bdf_b200.Engine(D, rank=rank, world=world)
        e1 = eng.add_entity_partitioned(N, rng.integers(0, world, N).astype(np.int32))
        e2 = eng.add_entity(M)
        rel = eng.add_relation([e1, e2], ids, vals)
        eng.set_relation_params(rel, 1.0, 0.4)
        eng.set_factors(e1, U)
        eng.set_factors(e2, V)
        assert rel_err(eng.predict_all(rel, (N, M)), want) <= 1e-13
        eng.close()
