"""The reference's own driver-level tests for this path, ported one to one onto the Python mirror of `macau()`:
test/alpha_sampling.jl, test/lambda_sampling.jl, test/rel_feat.jl, test/beta_saving.jl (they assert that the run completes,
the shapes of what it leaves behind and the dump file names — there is no asserted RMSE anywhere in the reference)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def dense_table(rng, n1, n2, k=2):
    A, B = rng.standard_normal((n1, k)), rng.standard_normal((n2, k))
    X = A @ B.T
    ii, jj = np.meshgrid(np.arange(1, n1 + 1), np.arange(1, n2 + 1), indexing="ij")
    ids = np.stack([ii.ravel(), jj.ravel()], 1).astype(np.int64)   # for i = 1:size(A,1), j = 1:size(B,1): push!(df, [i, j, X[i,j]])
    return ids, X.ravel().copy()


def test_alpha_sampling_jl():
    import bdf_b200

    Y = sp.random(15, 10, 0.1, random_state=1, format="csc")
    rd = bdf_b200.RelationData(Y, class_cut=0.5, alpha_sample=True)
    bdf_b200.assignToTest(rd.relations[0], 2, np.random.default_rng(0))
    bdf_b200.macau(rd, burnin=5, psamples=6, verbose=False)
    assert rd.relations[0].model.alpha > 0


def test_lambda_sampling_jl():
    import bdf_b200

    rng = np.random.default_rng(2)
    ids, vals = dense_table(rng, 30, 40)
    rd = bdf_b200.RelationData((ids, vals, [30, 40]))
    rd.entities[0].F = rng.standard_normal((30, 2))
    rd.entities[0].lambda_beta_sample = True
    bdf_b200.assignToTest(rd.relations[0], 10, rng)
    res = bdf_b200.macau(rd, burnin=50, psamples=10, num_latent=2, verbose=False)
    assert rd.entities[0].lambda_beta > 0 and np.isfinite(res["RMSE"])


def test_rel_feat_jl():
    import bdf_b200

    rng = np.random.default_rng(3)
    ids, vals = dense_table(rng, 30, 40)
    feat = rng.standard_normal((len(vals), 2))
    vals = vals + feat @ np.array([1.0, -1.0])
    rd = bdf_b200.RelationData((ids, vals, [30, 40]))
    rd.relations[0].model.alpha_sample = True
    rd.relations[0].F = feat
    bdf_b200.assignToTest(rd.relations[0], 10, rng)
    assert rd.relations[0].test_F.shape == (10, 2)
    res = bdf_b200.macau(rd, burnin=50, psamples=10, num_latent=2, verbose=False)
    assert np.isfinite(res["RMSE"]) and rd.relations[0].model.beta.shape == (2,)


def test_beta_saving_jl(tmp_path):
    import bdf_b200
    from bdf_b200 import data_reading as dr

    rng = np.random.default_rng(4)
    ids, vals = dense_table(rng, 20, 30)
    rd = bdf_b200.RelationData((ids, vals, [20, 30]), names=["A", "B"])
    rd.entities[0].F = rng.standard_normal((20, 3))
    bdf_b200.assignToTest(rd.relations[0], 10, rng)
    out = str(tmp_path / "macau-betasaving")
    bdf_b200.macau(rd, burnin=5, psamples=10, num_latent=2, verbose=False, output_beta=True, output=out, output_type="binary")
    for f in ("-A-01.binary", "-A-01.beta.binary", "-A-02.beta.binary"):
        assert os.path.isfile(out + f)
    b1 = dr.read_binary_float32(out + "-A-01.beta.binary")
    assert b1.shape == (3, 2)   # features × latents
    b10 = dr.read_binary_float32(out + "-A-10.beta.binary")
    assert np.allclose(b10, rd.entities[0].model.beta.astype(np.float32), rtol=1e-6, atol=1e-7)


def test_custom_rd_jl_macau_part():
    """test/custom_rd.jl:32-41: macau on hand-built entities with dense features, and on RelationData(r) from a sparse matrix."""
    import bdf_b200
    from bdf_b200.relation_data import Entity, Relation

    rng = np.random.default_rng(0)
    genes, pheno = Entity("genes"), Entity("pheno")
    genes.F = rng.random((100, 5))
    genes.lambda_beta = 3.0
    pheno.F = rng.random((50, 8))
    ids = np.stack([rng.integers(1, 101, 1050), rng.integers(1, 51, 1050)], 1)
    ids[0] = [100, 50]
    rd = bdf_b200.RelationData()
    rd.addRelation(Relation((ids, rng.random(1050)), "HPO", [genes, pheno], class_cut=0.5))
    res = bdf_b200.macau(rd, burnin=10, psamples=10, verbose=False)
    assert genes.model.beta.shape == (5, 10) and pheno.model.beta.shape == (8, 10) and res["gpu_launches"] > 0
    r2 = Relation(sp.random(100, 50, 0.01, random_state=1, format="csc"), "HPO2", [Entity("genes2"), Entity("pheno2")])
    bdf_b200.macau(bdf_b200.RelationData(r2), burnin=2, psamples=2, verbose=False)


def test_parallel_latent_tensor_jl():
    """test/parallel_latent_tensor.jl: 15 × 20 × 2 rank-2 tensor; entity names from the table columns; pred_all of a tensor equals
    the product of the three latent vectors of a cell plus the mean (:33-39)."""
    import bdf_b200

    rng = np.random.default_rng(6)
    A, B, Cc = rng.standard_normal((15, 2)), rng.standard_normal((20, 2)), rng.standard_normal((2, 2))
    X = np.einsum("id,jd,kd->ijk", A, B, Cc)
    ii, jj, kk = np.meshgrid(np.arange(1, 16), np.arange(1, 21), np.arange(1, 3), indexing="ij")
    ids = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], 1).astype(np.int64)
    rd = bdf_b200.RelationData((ids, X.ravel().copy(), [15, 20, 2]), names=["A", "B", "C"])
    assert [e.name for e in rd.entities] == ["A", "B", "C"]
    bdf_b200.assignToTest(rd.relations[0], 10, rng)
    eng = bdf_b200.Engine(2)
    res = bdf_b200.macau(rd, burnin=50, psamples=10, num_latent=2, verbose=False, engine=eng)
    assert np.isfinite(res["RMSE"])
    Yhat = eng.predict_all(0, (15, 20, 2))
    assert Yhat.shape == X.shape
    s = [e.model.sample for e in rd.entities]
    want = float(np.sum(s[0][3] * s[1][1] * s[2][0])) + rd.relations[0].model.mean_value   # Yhat[4,2,1]
    assert abs(Yhat[3, 1, 0] - want) <= 1e-12 * max(1.0, abs(want))
    eng.close()
